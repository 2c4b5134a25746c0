"""Device BVH builder vs host builder (development tool; results land in gpurun_out/build_bench.json and are summarised in
DESIGN.md): construction time and the traversal throughput the resulting tree gives, on the C3 and C5 scenes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

out = {}
for name, sc, w, h, spp in (("c3", scenes.c3_scene(), 1920, 1080, 8),) + ((("c5", scenes.c5_scene(), 3840, 2160, 2),) if os.environ.get("C5", "1") == "1" else ()):
    res = {"triangles": int(len(sc["tri9"]))}
    for kind in os.environ.get("KINDS", "device host").split():
        rt.BVH(sc["tri9"][:1000], on_device=(kind == "device"))          # warm up (CUDA context, OpenMP pool)
        t0 = time.time(); bvh = rt.BVH(sc["tri9"], on_device=(kind == "device")); wall = time.time() - t0
        info = bvh.info()
        scene = rt.Scene(sc["tri9"], sc["mat_idx"], sc["mats10"], sc["emissive"], skysphere=sc["env"], bvh=bvh)
        scene.trace_primary(sc["camera"], w, h)             # warm-up: the first launch of a kernel pays its lazy module load
        _, _, stp = scene.trace_primary(sc["camera"], w, h)
        scene.render(sc["camera"], w, h, 1, 8)
        img, st = scene.render(sc["camera"], w, h, spp, 8)
        res[kind] = dict(build_wall_s=wall, build_s=info["build_seconds"], inner=info["n_inner_nodes"], wide=info["n_wide_nodes"], depth=info["max_depth"],
                         wide_depth=info["wide_max_depth"], sah=info["sah_cost"], primary_mrays_s=w * h / stp["kernel_ms"] / 1e3, mrays_s=st["rays"] / st["kernel_ms"] / 1e3,
                         rays=int(st["rays"]), image_crc=int(np.bitwise_xor.reduce(img.view(np.uint32).ravel())))
        print(name, kind, res[kind], flush=True)
        del scene, bvh
    if "device" in res and "host" in res:
        res["images_equal"] = res["device"]["image_crc"] == res["host"]["image_crc"] and res["device"]["rays"] == res["host"]["rays"]
    out[name] = res
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "build_bench.json"), "w"), indent=1)
