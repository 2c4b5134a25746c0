"""Development tool: kernel time of one frame (or of one rank's share of it) per integrator. C3 by default.

    SPP=16 WORLDS="1 8" INTEGRATORS="1 2" python tools/integrator_bench.py

WORLD > 1 emulates rank 0 of an N-GPU frame on one GPU (its interleaved tiles only): the per-rank kernel time bounds the N-GPU frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

spp = int(os.environ.get("SPP", "16"))
worlds = [int(v) for v in os.environ.get("WORLDS", "1 8").split()]
integrators = [int(v) for v in os.environ.get("INTEGRATORS", "1 2").split()]
workload = os.environ.get("WORKLOAD", "c3")
w, h = (3840, 2160) if workload == "c5" else (1920, 1080)
s = scenes.c5_scene() if workload == "c5" else scenes.c3_scene()
sc = rt.Scene(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"], skysphere=s["env"])
fb = rt.Image(w, h, pinned=True).pixels
out = []
for world in worlds:
    for integ in integrators:
        sc.render(s["camera"], w, h, 1, 8, framebuffer=fb, rank=0, world=world, integrator=integ)
        best = None
        for rep in range(int(os.environ.get("REPS", "3"))):
            fb[...] = (0, 0, 0, 1)
            _, st = sc.render(s["camera"], w, h, spp, 8, framebuffer=fb, rank=0, world=world, integrator=integ)
            best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
        r = dict(workload=workload, world=world, integrator=integ, spp=spp, kernel_ms=best["kernel_ms"], rays=best["rays"],
                 mrays_s=best["rays"] / best["kernel_ms"] / 1e3, launches=best["gpu_launches"], env={k: v for k, v in os.environ.items() if k.startswith("B200RT_")})
        out.append(r)
        print(json.dumps(r), flush=True)
