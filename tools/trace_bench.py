"""Traversal microbenchmark (development tool): coherent primary rays and incoherent secondary-like rays on the C2/C3
meshes, 7-plane vs 3-plane, kernel-only CUDA-event times."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

dev = torch.device("cuda", 0)
c3 = scenes.c3_scene(sky_w=64, sky_h=32)
sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"])
cam = c3["camera"]
w, h = 1920, 1080
prim = torch.empty((h, w), dtype=torch.int32, device=dev)
t = torch.empty((h, w), dtype=torch.float32, device=dev)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(n):
        st = fn()
        ms.append(st["kernel_ms"])
    return min(ms), sum(ms) / len(ms)

res = {}
for flags, name in ((4, "7plane binary"), (16, "3plane binary"), (64, "8-ary")):
    best, avg = timeit(lambda: sc.trace_primary_device(cam, w, h, prim.data_ptr(), t.data_ptr(), flags=flags, want_stats=True))
    res[f"c3_primary_{name}"] = dict(ms=best, mrays=w * h / best / 1e3)
# secondary-like rays: from primary hit points, uniform random directions in the upper hemisphere of +y-ish / any direction
sc.trace_primary_device(cam, w, h, prim.data_ptr(), t.data_ptr())
torch.cuda.synchronize()
# reconstruct camera rays on the host via the API (trace_rays gives hit points)
pn, tn, _ = sc.trace_primary(cam, w, h)
hit = pn >= 0
ys, xs = np.nonzero(hit)
# camera ray directions (host restatement for the tool only)
m = cam.view_matrix; fov = cam.fov_dist
xn = (xs.astype(np.float32) / w * 2 - 1) * (w / h); yn = ys.astype(np.float32) / h * 2 - 1
pw = np.stack([xn, yn, np.full_like(xn, fov)], 1) @ m[:3, :3].T + m[:3, 3]
o = m[:3, 3][None, :].repeat(len(xs), 0)
d = pw - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
p = o + d * tn[hit][:, None]
rng = np.random.default_rng(1)
for name, gen in (("incoherent_sphere", lambda n: rng.normal(size=(n, 3))), ("up_hemisphere", lambda n: np.abs(rng.normal(size=(n, 3))) * np.array([1, 1, 1]) * np.where(rng.random((n, 3)) < 0.5, 1, -1) * np.array([1, 0, 1]) + np.abs(rng.normal(size=(n, 3))) * np.array([0, 1, 0]))):
    nd = gen(len(p)).astype(np.float32); nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    rays = np.concatenate([p + 1e-3 * nd, nd], 1).astype(np.float32)
    reps = max(1, 4_000_000 // len(rays))
    rays = np.tile(rays, (reps, 1))
    perm = rng.permutation(len(rays))
    for order, rr in (("screen_order", rays), ("shuffled", rays[perm])):
        dr = torch.from_numpy(np.ascontiguousarray(rr)).to(dev)
        n = len(rr)
        dp = torch.empty(n, dtype=torch.int32, device=dev); dt = torch.empty(n, dtype=torch.float32, device=dev)
        for flags, fname in ((4, "7plane binary"), (16, "3plane binary"), (0, "8-ary")):
            for any_hit in (False, True):
                best, avg = timeit(lambda: sc.trace_rays_device(dr.data_ptr(), n, dp.data_ptr(), dt.data_ptr(), any_hit=any_hit, flags=flags, want_stats=True))
                res[f"c3_{name}_{order}_{fname}_{'any' if any_hit else 'closest'}"] = dict(ms=best, mrays=n / best / 1e3, n=n, hit_frac=float((dp >= (1 if any_hit else 0)).float().mean()))
for k, v in res.items():
    print(f"{k:60s} {v['ms']:9.3f} ms  {v['mrays']:9.1f} Mrays/s  {v.get('hit_frac', '')}")
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "trace_bench.json"), "w"), indent=1)
