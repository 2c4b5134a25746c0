"""Diagnostic: camera rays (k_primary, binary layout) and one incoherent batch (k_trace_rays, 8-ary layout) on the host-built and the
device-built tree of C3, to be run under ncu: instruction counts tell tree quality apart from memory-layout effects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes
c3 = scenes.c3_scene()
rng = np.random.default_rng(1)
n = 1 << 20
o = rng.normal(size=(n, 3)); o = 6.0 * o / np.linalg.norm(o, axis=1, keepdims=True)
t = rng.normal(size=(n, 3)) * 1.5
d = t - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([o, d], 1).astype(np.float32)
for kind in ("host", "device"):
    bvh = rt.BVH(c3["tri9"], on_device=(kind == "device"))
    sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"], bvh=bvh)
    rt.rng_stream(0 if kind == "host" else 1, 1, 1, 1)          # marker launch
    prim, tt, st = sc.trace_primary(c3["camera"], 1920, 1080)
    p2, t2, _ = sc.trace_rays(rays)
    print(kind, "primary ms", st["kernel_ms"], "hits", int((prim >= 0).sum()), "batch hits", int((p2 >= 0).sum()), bvh.info())
