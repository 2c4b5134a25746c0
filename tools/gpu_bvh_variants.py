"""Host-builder parameter sweep on C3 (development tool): leaf size and SAH bin count vs path-tracing throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes
c3 = scenes.c3_scene()
spp = int(os.environ.get("SPP", "16"))
for leaf, bins in ((3, 16), (2, 16), (1, 16), (3, 32), (3, 8)):
    bvh = rt.BVH(c3["tri9"], max_leaf_size=leaf, sah_bins=bins)
    info = bvh.info()
    sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"], bvh=bvh)
    sc.render(c3["camera"], 1920, 1080, 1, 8)
    best = 1e9
    for rep in range(2):
        img, st = sc.render(c3["camera"], 1920, 1080, spp, 8)
        best = min(best, st["kernel_ms"])
    print(f"leaf {leaf} bins {bins}: {st['rays'] / best / 1e3:.1f} Mrays/s  {best:.2f} ms  wide nodes {info['n_wide_nodes']} depth {info['wide_max_depth']} sah {info['sah_cost']:.2f} build {info['build_seconds']:.2f}s", flush=True)
    del sc, bvh
