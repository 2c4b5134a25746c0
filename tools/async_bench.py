"""Development tool: the wavefront integrator's barrier-free continuation (csrc/async.cu) against the pass-synchronous frame and
the per-warp tail, whole frames and one rank of 8 emulated on one GPU, over the continuation's knobs — all inside one process
(the knobs are environment variables the library reads per frame). One JSON line per configuration.

    SPP=64 python tools/async_bench.py [config-set ...]          # sets: base thresholds shaders groups
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes

spp = int(os.environ.get("SPP", "64"))
reps = int(os.environ.get("REPS", "2"))
workload = os.environ.get("WORKLOAD", "c3")
sets = sys.argv[1:] or ["base"]
w, h = (3840, 2160) if workload == "c5" else (1920, 1080)
s = {"c5": scenes.c5_scene, "c3k": scenes.c3_knot_scene if hasattr(scenes, "c3_knot_scene") else scenes.c3_scene}.get(workload, scenes.c3_scene)()
sc = rt.Scene(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"], skysphere=s["env"])
fb = rt.Image(w, h, pinned=True).pixels
KNOBS = ("B200RT_WF_DETACH_AT", "B200RT_WF_DETACH_SLOTS", "B200RT_WF_DETACH_CTAS", "B200RT_WF_ASYNC", "B200RT_WF_ASYNC_PCT", "B200RT_WF_ASYNC_CAP", "B200RT_WF_ASYNC_MIN", "B200RT_WF_ASYNC_SHADERS", "B200RT_WF_GROUPS",
         "B200RT_WF_TAIL_PCT", "B200RT_WF_TAIL_CAP", "B200RT_WF_TAIL_MIN")

configs = []
def add(name, world, flags=None, **env):
    # configurations without explicit flags are the async study path
    configs.append((name, world, rt.FLAG_WF_ASYNC if flags is None else flags, {("B200RT_" + k): str(v) for k, v in env.items()}))

for world in (1, 8):
    if "base" in sets:
        add("passes only", world, rt.FLAG_WF_PASSES_ONLY)
        add("warp tail (round-2 default)", world, 0)
        add("async default", world)
    if "key" in sets:
        add("warp tail (round-2 default)", world, 0)
        for sh in ("1/4", "1/5", "1/6"):
            add(f"async default thresholds, shaders {sh}", world, WF_ASYNC_SHADERS=sh)
            add(f"async pct 20 cap 60000, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_PCT=20, WF_ASYNC_CAP=60000, WF_ASYNC_MIN=0)
            add(f"async from the start, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_MIN=10000000)
    if "prof" in sets and world == int(os.environ.get("PROF_WORLD", "1")):
        add("async from the start, 1 group (profiling shape)", world, WF_ASYNC_MIN=10000000, WF_GROUPS=1, WF_ASYNC_SHADERS=os.environ.get("PROF_SHADERS", "1/4"))
    if "default" in sets:
        add("default (passes + per-warp tail)", world, 0)
    if "detach_quick" in sets:
        add("no detach", world, 0)
        for slots in (6144, 12288):
            for ctas in (1, 2):
                add(f"detach at 36 slots {slots} ctas/SM {ctas}", world, rt.FLAG_WF_DETACH, WF_DETACH_AT=36, WF_DETACH_SLOTS=slots, WF_DETACH_CTAS=ctas)
        add("all tail from the start (groups <= 10M)", world, 0, WF_TAIL_MIN=10000000)
    if "detach" in sets:
        add("no detach", world, 0)
        add("detach default", world, rt.FLAG_WF_DETACH)
        for at in (18, 36, 72):
            for slots in (6144, 12288, 24576, 49152):
                for ctas in (1, 2, 3):
                    add(f"detach at {at} slots {slots} ctas/SM {ctas}", world, rt.FLAG_WF_DETACH, WF_DETACH_AT=at, WF_DETACH_SLOTS=slots, WF_DETACH_CTAS=ctas)
    if "quick" in sets:
        add("warp tail (round-2 default)", world, 0)
        for sh in ("1/4", "1/3"):
            add(f"async default thresholds, shaders {sh}", world, WF_ASYNC_SHADERS=sh)
            add(f"async from the start, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_MIN=10000000)
            add(f"async from the start, 1 group, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_MIN=10000000, WF_GROUPS=1)
            add(f"async from the start, 2 groups, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_MIN=10000000, WF_GROUPS=2)
    if "thresholds" in sets:
        for pct, cap in ((20, 60000), (30, 300000), (50, 400000), (70, 700000), (100, 10000000)):
            add(f"async pct {pct} cap {cap}", world, WF_ASYNC_PCT=pct, WF_ASYNC_CAP=cap, WF_ASYNC_MIN=0 if pct < 100 else 10000000)
    if "shaders" in sets:
        for sh in ("1/4", "1/3", "3/8", "1/2"):
            add(f"async shaders {sh}", world, WF_ASYNC_SHADERS=sh)
            add(f"async from the start, shaders {sh}", world, WF_ASYNC_SHADERS=sh, WF_ASYNC_MIN=10000000)
    if "groups" in sets:
        for g in (1, 2, 3, 4, 6):
            add(f"async groups {g}", world, WF_GROUPS=g)
            add(f"async from the start, groups {g}", world, WF_GROUPS=g, WF_ASYNC_MIN=10000000)

first = {}
for name, world, flags, env in configs:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    sc.render(s["camera"], w, h, 1, 8, framebuffer=fb, rank=0, world=world, flags=flags)
    best = None
    for rep in range(reps):
        fb[...] = (0, 0, 0, 1)
        _, st = sc.render(s["camera"], w, h, spp, 8, framebuffer=fb, rank=0, world=world, flags=flags)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    key = (world,)
    img = fb.copy()
    if key not in first:
        first[key] = (img, best["rays"])
    same = bool(np.array_equal(img.view(np.uint32), first[key][0].view(np.uint32))) and best["rays"] == first[key][1]
    print(json.dumps(dict(workload=workload, config=name, world=world, spp=spp, kernel_ms=round(best["kernel_ms"], 2), mrays_s=round(best["rays"] / best["kernel_ms"] / 1e3, 1),
                          launches=best["gpu_launches"], rays=best["rays"], equal_to_first=same, env=env)), flush=True)
