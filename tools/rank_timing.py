"""Emulates one rank of an N-GPU frame on a single GPU (development tool): rank 0's interleaved tiles of a world of
WORLD (default 8), C3 at SPP (default 64). The per-rank kernel time is what bounds the N-GPU frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes
world = int(os.environ.get("WORLD", "8")); spp = int(os.environ.get("SPP", "64"))
ranks = [int(v) for v in os.environ.get("RANKS", "0").split()]          # RANKS="0 1 2 3 4 5 6 7": the balance of the tile map
c3 = scenes.c3_scene()
sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"])
fb = rt.Image(1920, 1080).pixels
for rank in ranks:
    sc.render(c3["camera"], 1920, 1080, 1, 8, framebuffer=fb, rank=rank, world=world)
    for rep in range(3 if len(ranks) == 1 else 2):
        img, st = sc.render(c3["camera"], 1920, 1080, spp, 8, framebuffer=fb, rank=rank, world=world)
        print("world", world, "rank", rank, "spp", spp, "kernel_ms %.2f" % st["kernel_ms"], "launches", st["gpu_launches"], "rays", st["rays"], file=sys.stderr)
