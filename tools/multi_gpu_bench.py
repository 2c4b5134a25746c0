"""The N-GPU path BEHIND the C ABI (b200rt_scene_create_multi: one process, one host thread per device, peer copies into device 0's
gather buffer): frame time of b200rt_render on N = 1, 2, ... devices, and the frame's bit-equality with the 1-device frame.

    NS="1 2 4 8" SPP=64 WORKLOAD=c3 python tools/multi_gpu_bench.py        (run with gpurun --gpus N)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

spp = int(os.environ.get("SPP", "64"))
workload = os.environ.get("WORKLOAD", "c3")
have = torch.cuda.device_count()
ns = [n for n in (int(v) for v in os.environ.get("NS", "1 2 4 8").split()) if n <= have]
w, h = (3840, 2160) if workload == "c5" else (1920, 1080)
s = scenes.c5_scene() if workload == "c5" else scenes.c3_scene()
ref = None
for n in ns:
    t0 = time.time()
    sc = rt.Scene(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"], skysphere=s["env"], devices=list(range(n)))
    t_scene = time.time() - t0
    fb = rt.Image(w, h, pinned=True)
    sc.render(s["camera"], w, h, min(spp, 4), 8, framebuffer=fb.pixels)
    best = None
    for rep in range(int(os.environ.get("REPS", "3"))):
        fb.pixels[...] = (0, 0, 0, 1)
        t0 = time.perf_counter()
        _, st = sc.render(s["camera"], w, h, spp, 8, framebuffer=fb.pixels)
        wall = (time.perf_counter() - t0) * 1e3
        if best is None or wall < best[0]:
            best = (wall, st)
    img = fb.pixels.copy()
    if ref is None:
        ref = img
    wall, st = best
    print(json.dumps(dict(workload=workload, spp=spp, devices=n, scene_create_s=t_scene, frame_wall_ms=wall, slowest_rank_kernel_ms=st["kernel_ms"],
                          rays=st["rays"], mrays_s_wall=st["rays"] / wall / 1e3, equal_to_first=bool(np.array_equal(ref.view(np.uint32), img.view(np.uint32))),
                          h2d=st["h2d_bytes"], d2h=st["d2h_bytes"])), flush=True)
    del sc
