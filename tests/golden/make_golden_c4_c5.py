"""Round-2 golden fixtures, generated from the COMPILED REFERENCE ITSELF (oracle/_ref/libref_oracle.so) like make_golden.py:

    python tests/golden/make_golden_c4_c5.py

  render_c4_r{005,100}_m{1,0}.npz   BASELINE config 4 (roughness / metalness sweep on the Dragon-class stand-in) at its two extremes,
                                    c3small geometry (nu=100, nv=50), 64x32 procedural sky, 96x54 x 4 spp x 8 bounces
  render_c5small.npz                BASELINE config 5's scene class: 8 baked instances of a displaced sphere on a jittered grid + ground
  rng_wrap.npz                      per-pixel RNG streams (render_kernel.cpp:77-82) at 4K x 1024 spp pixels whose x*y*spp wraps int32
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle.oracle import RefOracle  # noqa: E402
from sycl_ray_tracing_b200 import scenes  # noqa: E402

C5_SMALL = dict(n_instances=8, nu=60, nv=24, sky_w=64, sky_h=32)
WRAP_PIXELS = [(3000, 2000), (3839, 2159), (2048, 2048), (1, 1), (3001, 1431)]      # (x, y) of a 3840x2160 frame, spp = 1024


def main():
    o = RefOracle()
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)
    dragon = o.camera_preset("dragon")

    for rough, tag_r in ((0.05, "005"), (1.0, "100")):
        for metal in (1, 0):
            c = scenes.c3_scene(roughness=rough, metalness=float(metal), nu=100, nv=50, sky_w=64, sky_h=32)
            s = o.scene_from_arrays(c["tri9"], c["mat_idx"], c["mats10"], c["emissive"])
            s.set_env(c["env"])
            img, _ = s.render(dragon, 96, 54, 4, 8)
            save(f"render_c4_r{tag_r}_m{metal}.npz", image=img, w=96, h=54, spp=4, bounces=8, nu=100, nv=50, sky_w=64, sky_h=32,
                 roughness=rough, metalness=metal)
            print(f"c4 r={rough} m={metal}: mean {img[..., :3].mean():.6f} nan {int((~np.isfinite(img)).sum())}")

    c5 = scenes.c5_scene(**C5_SMALL)
    s = o.scene_from_arrays(c5["tri9"], c5["mat_idx"], c5["mats10"], c5["emissive"])
    s.set_env(c5["env"])
    img, _ = s.render(dragon, 96, 54, 4, 8)
    prim, t, _ = s.primary(dragon, 192, 108, mode=0)
    pb, tb, _ = s.primary(dragon, 192, 108, mode=1)
    assert np.array_equal(t.view(np.uint32), tb.view(np.uint32))
    save("render_c5small.npz", image=img, w=96, h=54, spp=4, bounces=8, prim=prim, prim_brute=pb, t=t, pw=192, ph=108,
         **{k: np.int32(v) for k, v in C5_SMALL.items()})
    print(f"c5small: {len(c5['tri9'])} triangles, mean {img[..., :3].mean():.6f}, primary hits {(prim >= 0).sum()}, ties {(prim != pb).sum()}")

    kat = {}
    for x, y in WRAP_PIXELS:
        seed = (31 + x * y * 1024) & 0xFFFFFFFF            # int arithmetic wraps (-fwrapv build of the reference, render_kernel.cpp:77)
        st, fl = o.xorshift(seed, 10, 16)
        kat[f"state_{x}_{y}"] = np.array([st], np.uint32)
        kat[f"floats_{x}_{y}"] = fl
    save("rng_wrap.npz", pixels=np.array(WRAP_PIXELS, np.int32), spp=np.int32(1024), **kat)


if __name__ == "__main__":
    main()
