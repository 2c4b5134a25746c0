import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sycl_ray_tracing_b200 as rt
g = np.load("tests/golden/render_cornell_env.npz"); gs = np.load("tests/golden/scenes.npz"); gc = np.load("tests/golden/cameras.npz")
sc = rt.Scene(gs["cornell_tri9"], gs["cornell_mat_idx"], gs["cornell_mats10"], gs["cornell_emissive"], skysphere=g["env"])
c = rt.Camera.from_array17(gc["cornell"])
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
w, h = 100, 70
ref, st = sc.render(c, w, h, 2, 4)
for integ, name in ((1, "wavefront"), (2, "persistent")):
    for rep in range(3):
        fb = rt.Image(w, h).pixels
        rays = 0
        per = []
        for rank in range(3):
            _, s = sc.render(c, w, h, 2, 4, framebuffer=fb, integrator=integ, rank=rank, world=3)
            rays += s["rays"]; per.append(s["rays"])
        d = (bits(ref) != bits(fb)).any(-1)
        tiles_x = (w + 15) // 16
        ys, xs = np.nonzero(d)
        owner = ((ys // 16) * tiles_x + xs // 16) % 3
        print(name, "rep", rep, "differing px", int(d.sum()), list(zip(xs.tolist(), ys.tolist(), owner.tolist()))[:6], "rays", rays, st["rays"], per)
        for x, y in list(zip(xs.tolist(), ys.tolist()))[:3]:
            print("    ", (x, y), ref[y, x], fb[y, x], "single-pixel entry:", sc.ray_trace_pixel(c, w, h, 2, 4, x, y))
one, s1 = sc.render(c, w, h, 2, 4, integrator=2)
print("persistent one-shot equal:", np.array_equal(bits(one), bits(ref)), s1["rays"], st["rays"])
for integ in (1, 2):
    acc = rt.Accumulator(sc, c, w, h, 16, 5)
    rays = [acc.add(n, integrator=integ)["rays"] for n in (1, 4, 3, 8)]
    full, sf = sc.render(c, w, h, 16, 5)
    print("accum integrator", integ, "rays", rays, sum(rays), sf["rays"], "image equal:", np.array_equal(bits(acc.resolve()), bits(full)))
    acc2 = rt.Accumulator(sc, c, w, h, 16, 5)
    r2 = acc2.add(16, integrator=integ)["rays"]
    print("   single chunk rays", r2, "image equal:", np.array_equal(bits(acc2.resolve()), bits(full)))
