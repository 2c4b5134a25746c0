"""Development tool: how long the detached barrier-free kernels and the passes take inside one frame (B200RT_FLAG_TIME_INLINE)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sycl_ray_tracing_b200 as rt
from sycl_ray_tracing_b200 import scenes
world = int(os.environ.get("WORLD", "8")); spp = int(os.environ.get("SPP", "64"))
c3 = scenes.c3_scene()
sc = rt.Scene(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"], skysphere=c3["env"])
fb = rt.Image(1920, 1080).pixels
for name, env, flags in [("no detach", {}, 0),
                         ("detach 6144 slots, 1 CTA/SM", {"B200RT_WF_DETACH_SLOTS": "6144", "B200RT_WF_DETACH_CTAS": "1"}, rt.FLAG_WF_DETACH),
                         ("detach 6144 slots, 2 CTA/SM", {"B200RT_WF_DETACH_SLOTS": "6144", "B200RT_WF_DETACH_CTAS": "2"}, rt.FLAG_WF_DETACH),
                         ("detach 1024 slots, 1 CTA/SM", {"B200RT_WF_DETACH_SLOTS": "1024", "B200RT_WF_DETACH_CTAS": "1"}, rt.FLAG_WF_DETACH),
                         ("detach 1024 slots, 4 CTA/SM", {"B200RT_WF_DETACH_SLOTS": "1024", "B200RT_WF_DETACH_CTAS": "4"}, rt.FLAG_WF_DETACH)]:
    for k in ("B200RT_WF_DETACH_SLOTS", "B200RT_WF_DETACH_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["B200RT_WF_DETACH_AT"] = "36"
    sc.render(c3["camera"], 1920, 1080, 1, 8, framebuffer=fb, rank=0, world=world, flags=flags)
    for rep in range(2):
        _, st = sc.render(c3["camera"], 1920, 1080, spp, 8, framebuffer=fb, rank=0, world=world, flags=flags | rt.FLAG_TIME_INLINE)
    print(json.dumps(dict(config=name, world=world, kernel_ms=round(st["kernel_ms"], 2), trace_union_ms=round(st["trace_union_ms"], 2), trace_launches=st["trace_launches"],
                          trace_sum_ms=round(st["trace_ms"], 2), shade_sum_ms=round(st["shade_ms"], 2), tail_launches=st["tail_launches"], tail_sum_ms=round(st["tail_ms"], 2))), flush=True)
