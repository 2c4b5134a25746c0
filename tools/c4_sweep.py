"""BASELINE config 4: roughness x metalness sweep on the Dragon-class scene (geometry stays resident: b200rt_scene_set_materials).

    python tools/c4_sweep.py                      throughput per point -> gpurun_out/r2_c4_sweep.json
    MODE=ncu ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none \\
        -k regex:"wf_|k_rng" --csv --log-file gpurun_out/r2_c4_ncu.csv python tools/c4_sweep.py
    python tools/c4_sweep.py --merge gpurun_out/r2_c4_sweep.json gpurun_out/r2_c4_ncu.csv profiles/r2_c4_sweep.json

Under ncu (MODE=ncu) every sweep point renders one 2-spp frame with one tile group and no tail kernel, preceded by a k_rng_stream marker
launch, so the CSV splits into points; warp execution efficiency of a point = thread instructions / (32 x warp instructions) summed
over its wf_trace_coop (resp. wf_shade) launches (SURVEY 8d asks for it per point)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
POINTS = [(m, r) for m in (1.0, 0.0) for r in (0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0)]

if len(sys.argv) > 1 and sys.argv[1] == "--merge":
    sweep = json.load(open(sys.argv[2]))
    rows = list(csv.reader(l for l in open(sys.argv[3]) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        launches.setdefault(int(r[iid]), {"kernel": r[ik].split("(")[0]})[r[im]] = float(r[iv].replace(",", ""))
    point = -1
    acc = {}
    for lid in sorted(launches):
        L = launches[lid]
        if L["kernel"].startswith("k_rng"):
            point += 1
            continue
        if point < 0 or "smsp__inst_executed.sum" not in L:
            continue
        a = acc.setdefault((point, L["kernel"]), [0.0, 0.0])
        a[0] += L["smsp__inst_executed.sum"] * L["smsp__thread_inst_executed_per_inst_executed.ratio"]
        a[1] += L["smsp__inst_executed.sum"]
    for i, p in enumerate(sweep["sweep"]):
        for k in ("wf_trace_coop", "wf_shade"):
            if (i, k) in acc and acc[(i, k)][1] > 0:
                p[f"warp_execution_efficiency_{k}"] = acc[(i, k)][0] / acc[(i, k)][1] / 32.0
                p[f"warp_instructions_{k}_2spp"] = acc[(i, k)][1]
    json.dump(sweep, open(sys.argv[4], "w"), indent=1)
    for p in sweep["sweep"]:
        print(p)
    sys.exit(0)

import numpy as np
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

ncu_mode = os.environ.get("MODE") == "ncu"
spp = 2 if ncu_mode else int(os.environ.get("SPP", "16"))
c3 = scenes.c3_scene()
sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"])
sweep = []
for i, (metal, rough) in enumerate(POINTS):
    m = c3["mats10"].copy(); m[1, 8] = metal; m[1, 9] = rough
    sc.set_materials(m)
    if ncu_mode:
        rt.rng_stream(i, 1, 1, 1)           # marker launch
        sc.render(c3["camera"], 1920, 1080, spp, 8)
        continue
    sc.render(c3["camera"], 1920, 1080, 1, 8)
    img, st = sc.render(c3["camera"], 1920, 1080, spp, 8)
    r = dict(metalness=metal, roughness=rough, mrays_s=st["rays"] / st["kernel_ms"] / 1e3, mspp_s=st["samples"] / st["kernel_ms"] / 1e3,
             rays_per_sample=st["rays"] / st["samples"], kernel_ms=st["kernel_ms"], nan_px=int(np.isnan(img[..., :3]).any(-1).sum()),
             mean=float(np.nanmean(img[..., :3])))
    sweep.append(r); print(r, flush=True)
if not ncu_mode:
    json.dump(dict(spp=spp, sweep=sweep), open(os.path.join(ROOT, "gpurun_out", "r2_c4_sweep.json"), "w"), indent=1)
