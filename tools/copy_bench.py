"""Development tool: host<->device copy overhead of b200rt_render / b200rt_render_rgba8 / b200rt_trace_primary with pageable and
page-locked caller buffers (total_ms - kernel_ms of a frame whose kernel time is small)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sycl_ray_tracing_b200 as rt

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes.npz"))
sc = rt.Scene(g["cornell_tri9"], g["cornell_mat_idx"], g["cornell_mats10"], g["cornell_emissive"], skysphere=rt.constant_env(1.0))
cam = rt.Camera.CORNELL_BOX_CAMERA
w, h = 1920, 1080
for name, fb in (("pageable", rt.Image(w, h).pixels), ("pinned", rt.Image(w, h, pinned=True).pixels)):
    for rep in range(4):
        fb[...] = (0.1, 0.1, 0.1, 1)
        t0 = time.perf_counter()
        _, st = sc.render(cam, w, h, 1, 2, framebuffer=fb)
        wall = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(api="render", buffers=name, wall_ms=wall, total_ms=st["total_ms"], kernel_ms=st["kernel_ms"], copy_ms=st["total_ms"] - st["kernel_ms"],
                          h2d=st["h2d_bytes"], d2h=st["d2h_bytes"])), flush=True)
out8 = np.empty((h, w, 4), np.uint8)
for rep in range(4):
    t0 = time.perf_counter()
    _, st = sc.render_rgba8(cam, w, h, 1, 2, out=out8)
    wall = (time.perf_counter() - t0) * 1e3
print(json.dumps(dict(api="render_rgba8", buffers="pageable", wall_ms=wall, total_ms=st["total_ms"], kernel_ms=st["kernel_ms"], d2h=st["d2h_bytes"])), flush=True)
for rep in range(4):
    t0 = time.perf_counter()
    prim, t, st = sc.trace_primary(cam, w, h)
    wall = (time.perf_counter() - t0) * 1e3
print(json.dumps(dict(api="trace_primary", buffers="pageable", wall_ms=wall, total_ms=st["total_ms"], kernel_ms=st["kernel_ms"])), flush=True)
