"""Generates the committed golden fixtures from the COMPILED REFERENCE ITSELF (oracle/_ref/libref_oracle.so, i.e. the
unmodified sources under /root/reference driven by oracle/ref_driver.cpp). Run in the build container:

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so everything the GPU tests need from it is frozen here:
  scenes.npz          Utils::parse_obj output for the three bundled OBJs (triangles, material indices, materials, emissive list)
  cameras.npz         the five Camera presets (16 matrix floats + fov_dist)
  bvh_tests.npz       include/bvh_tests.h: 572 hit rays + expected points, 222 miss rays, and BVH::intersect's prim/t for them
  xorshift.npz        xorshift32_generator known answers
  primary_*.npz       un-jittered primary-ray prim/t maps from BVH::intersect
  render_*.npz        RenderKernel::render() framebuffers (tone-mapped RGBA, as main.cpp sees them)
  env_cdf.npz         Utils::compute_env_map_cdf of a small procedural sky
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle.oracle import REF_DIR, RefOracle  # noqa: E402
from sycl_ray_tracing_b200 import scenes  # noqa: E402

OBJS = {"cornell": "cornell_pbr.obj", "mis": "MIS.obj", "area": "test_triangle_area_sampling.obj"}
CAMS = {"cornell": "cornell", "mis": "mis", "area": "cornell"}


def test_env(seed=5, w=32, h=16):
    rng = np.random.default_rng(seed)
    env = (rng.random((h, w, 4)) * 2.0).astype(np.float32)
    env[3:5, 7:9, :3] += 500.0           # a "sun"
    env[..., 3] = 0.0
    return env


def main():
    o = RefOracle()
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)

    cams = {n: o.camera_preset(n) for n in ["cornell", "ganesha", "ite_orb", "dragon", "mis"]}
    cams["c2"] = o.camera_make(45.0, 0.0, 0.0, (0.0, 0.0, 10.5))
    save("cameras.npz", **cams)

    sc = {}
    ref_scenes = {}
    for key, fn in OBJS.items():
        s = o.scene_from_obj(os.path.join(REF_DIR, "data", "OBJs", fn))
        a = o.scene_arrays(s)
        for k, v in a.items():
            sc[f"{key}_{k}"] = v
        ref_scenes[key] = s
    save("scenes.npz", **sc)

    hit, pts, miss = o.golden()
    cs = ref_scenes["cornell"]
    hp, ht, hex_ = cs.trace(hit, mode=0, extra=True)
    mp, mt = cs.trace(miss, mode=0)
    save("bvh_tests.npz", hit_rays=hit, hit_points=pts, miss_rays=miss, hit_prim=hp, hit_t=ht, hit_extra=hex_, miss_prim=mp, miss_t=mt)

    kat = {}
    for seed in (31, 591, 1, 0xFFFFFFFF, 31 + 3839 * 2159 * 1024 & 0xFFFFFFFF):
        st, fl = o.xorshift(seed, 10, 32)
        kat[f"state_{seed}"] = np.array([st], np.uint32)
        kat[f"floats_{seed}"] = fl
    save("xorshift.npz", **kat)

    # prim = the octree's answer (BVH::intersect), prim_brute = RenderKernel::intersect_scene's (lowest index on exact-t ties)
    for key, s in ref_scenes.items():
        prim, t, _ = s.primary(cams[CAMS[key]], 128, 128, mode=0)
        pb, tb, _ = s.primary(cams[CAMS[key]], 128, 128, mode=1)
        assert np.array_equal(t.view(np.uint32), tb.view(np.uint32))
        save(f"primary_{key}.npz", prim=prim.astype(np.int16), prim_brute=pb.astype(np.int16), t=t, w=128, h=128)

    env_const = np.full((2, 4, 4), 1.0e-20, np.float32); env_const[..., 3] = 0.0
    env_test = test_env()
    renders = [("cornell", env_const, 64, 64, 8, 4, "c1"), ("cornell", env_test, 64, 64, 4, 6, "env"),
               ("mis", env_test, 64, 64, 4, 4, "env"), ("area", env_const, 64, 64, 4, 3, "c1")]
    for key, env, w, h, spp, b, tag in renders:
        s = ref_scenes[key]
        s.set_env(env)
        img, _ = s.render(cams[CAMS[key]], w, h, spp, b)
        save(f"render_{key}_{tag}.npz", image=img, env=env, w=w, h=h, spp=spp, bounces=b, cdf=s.env_cdf())

    # C2-class (small): displaced sphere, primary hits from the reference octree
    c2 = scenes.c2_scene(nu=100, nv=50)
    s = o.scene_from_arrays(c2["tri9"], c2["mat_idx"], c2["mats10"], c2["emissive"])
    prim, t, _ = s.primary(cams["c2"], 240, 135, mode=0)
    pb, tb, _ = s.primary(cams["c2"], 240, 135, mode=1)
    assert np.array_equal(t.view(np.uint32), tb.view(np.uint32))
    print("c2small: exact-t ties where the octree's traversal order and brute force pick different triangles:", int((prim != pb).sum()))
    save("primary_c2small.npz", prim=prim.astype(np.int32), prim_brute=pb.astype(np.int32), t=t, w=240, h=135, nu=100, nv=50)

    # C3-class (small): metal displaced sphere on a ground quad under the procedural sun+sky
    c3 = scenes.c3_scene(roughness=0.25, nu=100, nv=50, sky_w=64, sky_h=32)
    s = o.scene_from_arrays(c3["tri9"], c3["mat_idx"], c3["mats10"], c3["emissive"])
    s.set_env(c3["env"])
    img, _ = s.render(cams["dragon"], 96, 54, 4, 8)
    save("render_c3small.npz", image=img, w=96, h=54, spp=4, bounces=8, nu=100, nv=50, sky_w=64, sky_h=32, roughness=0.25)
    save("env_cdf.npz", cdf=s.env_cdf(), sky_w=64, sky_h=32)

    # spheres: cornell + one analytic metal sphere (the variant commented out at main.cpp:74-75)
    a = o.scene_arrays(ref_scenes["cornell"])
    mats = np.concatenate([a["mats10"], np.array([[0, 0, 0, 1, 1.0, 0.71, 0.29, 1, 1.0, 0.4]], np.float32)])
    sph = np.array([[0.3275, 0.7, 0.3725, 0.2]], np.float32)
    s = o.scene_from_arrays(a["tri9"], a["mat_idx"], mats, a["emissive"], spheres4=sph, sphere_prim=np.array([len(a["tri9"])], np.int32),
                            sphere_mat_idx=np.array([len(mats) - 1], np.int32))
    s.set_env(env_test)
    prim, t, _ = s.primary(cams["cornell"], 128, 128, mode=3)
    img, _ = s.render(cams["cornell"], 64, 64, 4, 4)
    save("sphere_cornell.npz", prim=prim.astype(np.int16), t=t, image=img, env=env_test, spheres4=sph, mats10=mats, w=64, h=64, spp=4, bounces=4)

    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"golden fixtures written to {HERE}: {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
