"""BASELINE configs 4 and 5 (development tool; results land in gpurun_out/configs.json and are summarised in DESIGN.md):
C4 = metal/rough sweep on the Dragon-class scene (re-uses the resident geometry, b200rt_scene_set_materials);
C5 = 20 M-triangle procedural scene at 3840x2160 (reduced spp: the full 1024 spp frame is ~65 G rays)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

out = {}
spp = int(os.environ.get("SPP", "16"))
c3 = scenes.c3_scene()
sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"])
sweep = []
for metal in (1.0, 0.0):
    for rough in (0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0):
        m = c3["mats10"].copy(); m[1, 8] = metal; m[1, 9] = rough
        sc.set_materials(m)
        sc.render(c3["camera"], 1920, 1080, 1, 8)
        img, st = sc.render(c3["camera"], 1920, 1080, spp, 8)
        r = dict(metalness=metal, roughness=rough, mrays_s=st["rays"] / st["kernel_ms"] / 1e3, mspp_s=st["samples"] / st["kernel_ms"] / 1e3,
                 rays_per_sample=st["rays"] / st["samples"], kernel_ms=st["kernel_ms"], nan_px=int(np.isnan(img[..., :3]).any(-1).sum()),
                 mean=float(np.nanmean(img[..., :3])))
        sweep.append(r); print(r, flush=True)
out["c4"] = dict(spp=spp, sweep=sweep)
del sc
if os.environ.get("C5", "1") == "1":
    t0 = time.time(); c5 = scenes.c5_scene(); t_gen = time.time() - t0
    t0 = time.time(); s5 = rt.Scene(c5["tri9"], c5["mat_idx"], c5["mats10"], c5["emissive"], skysphere=c5["env"]); t_scene = time.time() - t0
    info = s5.bvh_info()
    prim, t, stp = s5.trace_primary(c5["camera"], 3840, 2160)
    s5.render(c5["camera"], 3840, 2160, 1, 8)
    img, st = s5.render(c5["camera"], 3840, 2160, 4, 8)
    img2, st2 = s5.render(c5["camera"], 3840, 2160, 4, 8, flags=rt.FLAG_BVH2)          # ablation: binary layout on the HBM-bound scene
    _, _, stp2 = s5.trace_primary(c5["camera"], 3840, 2160, flags=rt.FLAG_BVH2)
    out["c5"] = dict(triangles=len(c5["tri9"]), gen_s=t_gen, scene_create_s=t_scene, bvh=info, device_bytes=s5.device_bytes(),
                     primary_mrays_s=3840 * 2160 / stp["kernel_ms"] / 1e3, primary_hits=int((prim >= 0).sum()),
                     spp=4, mrays_s=st["rays"] / st["kernel_ms"] / 1e3, mspp_s=st["samples"] / st["kernel_ms"] / 1e3,
                     bvh2_mrays_s=st2["rays"] / st2["kernel_ms"] / 1e3, bvh2_primary_mrays_s=3840 * 2160 / stp2["kernel_ms"] / 1e3,
                     bvh2_image_equal=bool(np.array_equal(img.view(np.uint32), img2.view(np.uint32))),
                     rays_per_sample=st["rays"] / st["samples"], kernel_ms=st["kernel_ms"], bytes_per_ray=scenes.algorithmic_bytes_per_ray(len(c5["tri9"])))
    print(out["c5"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
