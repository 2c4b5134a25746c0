"""Diagnostic: is the device BVH builder deterministic (two builds of the 20 M-triangle scene: identical node / triangle arrays?), and is a
frame's ray count (same scene object, two renders; two scene objects)?"""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes
c5 = scenes.c5_scene()
hs = []
scs = []
for i in range(2):
    bvh = rt.BVH(c5["tri9"], on_device=True)
    f = bvh.flatten()
    h = (hashlib.md5(f.wide.tobytes()).hexdigest(), hashlib.md5(np.ascontiguousarray(f.tris).tobytes()).hexdigest(), hashlib.md5(np.ascontiguousarray(f.axis).tobytes()).hexdigest())
    print("build", i, bvh.info()["n_wide_nodes"], h, flush=True)
    hs.append(h)
    scs.append(rt.Scene(c5["tri9"], c5["mat_idx"], c5["mats10"], c5["emissive"], skysphere=c5["env"], bvh=bvh))
    del f
print("device builds identical:", hs[0] == hs[1])
spp = int(os.environ.get("SPP", "32"))
res = []
for name, sc in (("scene0", scs[0]), ("scene0 again", scs[0]), ("scene1", scs[1])):
    img, st = sc.render(c5["camera"], 3840, 2160, spp, 8)
    res.append((st["rays"], hashlib.md5(img.tobytes()).hexdigest()))
    print(name, "rays", st["rays"], "image md5", res[-1][1], flush=True)
print("ray counts equal:", len({r[0] for r in res}) == 1, "images equal:", len({r[1] for r in res}) == 1)
m, stm = scs[0].render(c5["camera"], 3840, 2160, spp, 8, integrator=rt.INTEGRATOR_MEGAKERNEL)
print("megakernel rays", stm["rays"], "image md5", hashlib.md5(m.tobytes()).hexdigest())
