"""not gpu: the C-ABI library loads and exports what include/b200rt.h declares; host-side logic (BVH builder, tile
bookkeeping, Camera/Image/CDF mirrors, scene generators). No CUDA compute is invoked."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda, load_golden, scene_arrays


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_library_exports_every_declared_symbol(rt):
    header = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    declared = sorted(set(re.findall(r"\b(b200rt_[a-z_0-9]+)\s*\(", header)))
    assert set(declared) == set(rt.binding.EXPORTS), set(declared) ^ set(rt.binding.EXPORTS)
    lib = rt.load_library()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.b200rt_version()


def test_header_cites_reference_interfaces():
    header = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    for cite in ("render_kernel.h:24-46", "render_kernel.cpp:189-211", "flattened_bvh.h:25-39", "bvh.cpp:62-65", "utils.cpp:126-142",
                 "image_io.cpp:165-182", "bvh.cpp:19-60", "render_kernel.cpp:532-567"):
        assert cite in header


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu(rt, golden_scenes):
    a = scene_arrays(golden_scenes, "cornell")
    with pytest.raises(rt.B200RTError, match="no CUDA device"):
        rt.Scene(a["tri9"], a["mat_idx"], a["mats10"], a["emissive"])
    # the stages either side of the path have no CPU fallback either
    with pytest.raises(rt.B200RTError, match="no CUDA device"):
        rt.BVH(a["tri9"], on_device=True)
    with pytest.raises(rt.B200RTError, match="no CUDA device"):
        rt.quantise_rgba8(np.zeros((4, 4, 4), np.float32))


@pytest.mark.parametrize("key", ["cornell", "mis", "area"])
def test_bvh_builder_invariants_bundled(rt, golden_scenes, key):
    a = scene_arrays(golden_scenes, key)
    for leaf, diag in ((4, True), (1, True), (15, False)):
        b = rt.BVH(a["tri9"], max_leaf_size=leaf, use_diag_slabs=diag)
        b.check()
        info = b.info()
        assert info["n_triangles"] == len(a["tri9"]) and info["max_leaf_size"] <= leaf and info["max_depth"] <= 60
        assert info["n_leaves"] == info["n_inner_nodes"] + 1 or info["n_inner_nodes"] == 1
        f = b.flatten()
        assert f.axis.shape == (info["n_inner_nodes"], 16) and f.tris.shape == (len(a["tri9"]), 12)
        assert sorted(f.tris[:, 3].view(np.int32).tolist()) == list(range(len(a["tri9"])))


def test_bvh_builder_degenerate_inputs(rt):
    rng = np.random.default_rng(3)
    cases = {
        "empty": np.zeros((0, 9), np.float32),
        "one": rng.random((1, 9)).astype(np.float32),
        "two": rng.random((2, 9)).astype(np.float32),
        "duplicates": np.repeat(rng.random((1, 9)).astype(np.float32), 300, axis=0),
        "collinear_centroids": np.concatenate([np.stack([np.full(9, i, np.float32) + np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32) * 0 for i in range(100)])]),
        "huge_and_tiny": np.concatenate([rng.random((50, 9)).astype(np.float32) * 1e6, rng.random((50, 9)).astype(np.float32) * 1e-6]),
        "random_soup": (rng.random((5000, 9)).astype(np.float32) * 10 - 5),
    }
    for name, tri in cases.items():
        b = rt.BVH(tri)
        b.check()
        assert b.info()["max_depth"] <= 60, name


def test_bvh_builder_c2_mesh(rt):
    from sycl_ray_tracing_b200 import scenes
    c2 = scenes.c2_scene(nu=200, nv=100)
    assert c2["tri9"].shape == (40000, 9)
    b = rt.BVH(c2["tri9"])
    b.check()
    info = b.info()
    assert info["max_depth"] <= 40 and info["sah_cost"] < 200


def test_leaf_reference_packing(rt, golden_scenes):
    a = scene_arrays(golden_scenes, "mis")
    f = rt.BVH(a["tri9"]).flatten()
    refs = f.axis[:, 12:14].view(np.int32)
    counts = f.axis[:, 14:16].view(np.int32)
    leaf = refs < 0
    packed = ~refs[leaf]
    assert np.array_equal(packed & 15, counts[leaf]), "count lives in the low 4 bits of ~ref"
    first = packed >> 4
    assert (first >= 0).all() and (first + counts[leaf] <= len(a["tri9"])).all()
    assert counts[leaf].sum() == len(a["tri9"]) and (counts[~leaf] == 0).all()
    assert (refs[~leaf] > 0).all() and (refs[~leaf] < len(f.axis)).all()


def test_tiles_for_rank_partition(rt):
    from sycl_ray_tracing_b200 import distributed as D
    for (w, h) in [(1920, 1080), (512, 512), (100, 70), (16, 16), (17, 1), (3840, 2160)]:
        n = ((w + 15) // 16) * ((h + 15) // 16)
        for world in (1, 2, 3, 4, 8):
            per = [rt.tiles_for_rank(w, h, r, world) for r in range(world)]
            assert per == [D.tiles_for_rank(w, h, r, world) for r in range(world)]
            assert sum(per) == n and max(per) == per[0] and max(per) - min(per) <= 1
            seen = np.zeros((h, w), np.int32)
            for r in range(world):
                xy = D.tile_slot_coords(w, h, r, world, per[0])
                ok = xy[:, 0] >= 0
                np.add.at(seen, (xy[ok, 1], xy[ok, 0]), 1)
            assert (seen == 1).all(), "every pixel belongs to exactly one rank's tile buffer"
    assert rt.tiles_for_rank(0, 10, 0, 1) == 0 and rt.tiles_for_rank(10, 10, 2, 2) == 0


def test_camera_presets_and_image_mirror(rt, golden_cameras):
    for name, key in [("CORNELL_BOX_CAMERA", "cornell"), ("GANESHA_CAMERA", "ganesha"), ("ITE_ORB_CAMERA", "ite_orb"),
                      ("PBRT_DRAGON_CAMERA", "dragon"), ("MIS_CAMERA", "mis")]:
        assert np.array_equal(bits(getattr(rt.Camera, name).as_array17()), bits(golden_cameras[key])), name
    c = rt.Camera(45.0, rt.Translation(0.0, 0.0, 10.5))
    assert np.array_equal(bits(c.as_array17()), bits(golden_cameras["c2"]))
    # a rotated camera agrees with the reference's sinf/cosf to an ulp
    c = rt.Camera(45.0, rt.api.compose_transform(rt.RotationX(-45.0), rt.Translation(0.0, -1.0, 10.5)))
    assert np.allclose(c.as_array17(), golden_cameras["dragon"], rtol=3e-7, atol=1e-7)
    img = rt.Image(5, 3)
    assert img.width() == 5 and img.height() == 3 and np.array_equal(img.pixels[0, 0], [0, 0, 0, 1])
    m = rt.SimpleMaterial()
    assert np.array_equal(m.as_array10(), np.array([0, 0, 0, 1, 1, 0.2, 0.7, 1, 0, 1], np.float32))


def test_env_cdf_mirror_matches_reference(rt):
    from sycl_ray_tracing_b200 import scenes
    g = load_golden("env_cdf.npz")
    sky = scenes.procedural_sky(int(g["sky_w"]), int(g["sky_h"]))
    assert np.array_equal(bits(rt.compute_env_map_cdf(sky)), bits(g["cdf"]))
    r = load_golden("render_cornell_env.npz")
    assert np.array_equal(bits(rt.compute_env_map_cdf(r["env"])), bits(r["cdf"]))


def test_scene_generators_are_deterministic_and_sized(rt):
    from sycl_ray_tracing_b200 import scenes
    a, b = scenes.displaced_sphere(64, 32), scenes.displaced_sphere(64, 32)
    assert a.shape == (2 * 64 * 32, 9) and np.array_equal(bits(a), bits(b))
    c3 = scenes.c3_scene(nu=64, nv=32, sky_w=32, sky_h=16)
    assert len(c3["tri9"]) == 2 * 64 * 32 + 2 and c3["mat_idx"][-1] == 2 and c3["env"].shape == (16, 32, 4)
    assert c3["env"][..., :3].max() > 100 and (c3["env"][..., 3] == 0).all()
    out, inn = scenes.displaced_sphere(16, 8, outward=True), scenes.displaced_sphere(16, 8, outward=False)
    n_out = np.cross(out[:, 3:6] - out[:, :3], out[:, 6:9] - out[:, :3])
    assert (np.einsum("ij,ij->i", n_out, out[:, :3]) > 0).mean() > 0.95, "outward winding faces away from the centre"
    assert not np.array_equal(out, inn)
    assert [scenes.algorithmic_bytes_per_ray(n) for n in (32, 10 ** 6, 870000, 2 * 10 ** 7)] == [768, 3008, 3008, 3904]


def test_c2_known_answer_through_the_oracle():
    """502 161 of the 2 073 600 un-jittered primary rays hit the 1M-triangle C2 mesh."""
    from oracle.oracle import PortOracle
    from sycl_ray_tracing_b200 import scenes
    c2 = scenes.c2_scene()
    ps = PortOracle().scene_from_arrays(c2["tri9"], c2["mat_idx"], c2["mats10"], c2["emissive"])
    prim, t, _ = ps.primary(c2["camera"].as_array17(), 1920, 1080, mode=0)
    assert int((prim >= 0).sum()) == 502161
