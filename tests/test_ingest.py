"""Host-side ingest of the C ABI (b200rt_obj_load, b200rt_hdr_load): CPU tests, no GPU needed.

The OBJ reader is checked against the reference's own parse of its three bundled scenes (tests/golden/scenes.npz, frozen from
Utils::parse_obj by make_golden.py) when /root/reference is present; the HDR reader against the reference's stb round trip
through oracle/_ref, and always against a file written here from the format description."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_golden

REF_OBJS = "/root/reference/data/OBJs"
OBJS = {"cornell": "cornell_pbr.obj", "mis": "MIS.obj", "area": "test_triangle_area_sampling.obj"}


@pytest.mark.parametrize("key", list(OBJS))
@pytest.mark.skipif(not os.path.isdir(REF_OBJS), reason="the reference's bundled OBJ files are only present in the build container")
def test_parse_obj_equals_the_reference_parse(rt, key):
    g = load_golden("scenes.npz")
    o = rt.parse_obj(os.path.join(REF_OBJS, OBJS[key]))
    assert np.array_equal(o["tri9"].view(np.uint32), g[f"{key}_tri9"].view(np.uint32)), "triangle list (order, quad diagonals) must match"
    assert np.array_equal(o["mat_idx"], g[f"{key}_mat_idx"]) and np.array_equal(o["emissive"], g[f"{key}_emissive"])
    assert np.array_equal(o["mats10"].view(np.uint32), g[f"{key}_mats10"].view(np.uint32))


def test_parse_obj_rules(rt, tmp_path):
    """A hand-written OBJ: negative indices, v/vt/vn references, a quad cut along its shorter diagonal (both cases), a pentagon
    fanned, faces before any usemtl -> default material 0, roughness clamp, illum 0 rule, emissive list, errors."""
    (tmp_path / "m.mtl").write_text("newmtl light\nKe 2 2 2\nKd 0 0 0\nPr 0.5\nPm 0.25\nillum 2\n\nnewmtl glossy\nKd 0.1 0.2 0.3\nPr 0\nPm 1\nillum 2\n"
                                   "newmtl legacy\nKd 0.5 0.5 0.5\nPr 0.3\nPm 0.7\n")
    (tmp_path / "s.obj").write_text(
        "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 3 0 0\nv 0 0.5 0\nv 2 2 1\n"
        "f 1 2 3\n"                       # default material
        "usemtl light\nf 1/1/1 2/2/2 3/3/3 4/4/4\n"          # square: d02 == d13 -> 1-3 diagonal
        "usemtl glossy\nf 1 5 3 6\n"      # |p0-p2|^2 = 2 < |p1-p3|^2 = 9.25 -> 0-2 diagonal
        "usemtl legacy\nf -7 -6 -5 -4 -1\n")
    o = rt.parse_obj(str(tmp_path / "s.obj"))
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [3, 0, 0], [0, 0.5, 0], [2, 2, 1]], np.float32)
    T = lambda a, b, c: np.concatenate([V[a], V[b], V[c]])
    want = np.stack([T(0, 1, 2), T(0, 1, 3), T(1, 2, 3), T(0, 4, 2), T(0, 2, 5), T(0, 1, 2), T(0, 2, 3), T(0, 3, 6)])
    assert np.array_equal(o["tri9"], want)
    assert o["mat_idx"].tolist() == [0, 1, 1, 2, 2, 3, 3, 3] and o["emissive"].tolist() == [1, 2]
    m = o["mats10"]
    assert m.shape == (4, 10) and m[0].tolist() == [1, 0, 1, 1, 0, 0, 0, 1, 0, 1]
    assert m[1].tolist() == [2, 2, 2, 1, 0, 0, 0, 1, 0.25, 0.5]
    assert np.allclose(m[2], [0, 0, 0, 1, 0.1, 0.2, 0.3, 1, 1.0, 0.01])            # roughness 0 clamped to 1e-2 (utils.cpp:82)
    assert np.allclose(m[3], [0, 0, 0, 1, 0.5, 0.5, 0.5, 1, 0.0, 1.0])             # no illum line = illum 0: default roughness / metalness
    with pytest.raises(rt.B200RTError):
        rt.parse_obj(str(tmp_path / "missing.obj"))
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(rt.B200RTError):
        rt.parse_obj(str(tmp_path / "bad.obj"))
    (tmp_path / "nomtl.obj").write_text("mtllib nothere.mtl\nv 0 0 0\n")
    with pytest.raises(rt.B200RTError):
        rt.parse_obj(str(tmp_path / "nomtl.obj"))


def write_flat_hdr(path, rgb):
    h, w, _ = rgb.shape
    m = rgb.max(axis=-1)
    mant, ex = np.frexp(m.astype(np.float64))
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, ex + 128, 0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())
    dec = np.where(rgbe[..., 3:4] != 0, rgbe[..., :3].astype(np.float32) * np.ldexp(np.float32(1.0), rgbe[..., 3:4].astype(np.int32) - 136), 0.0)
    return dec.astype(np.float32)


def test_hdr_reader_flat_scanlines(rt, tmp_path):
    rng = np.random.default_rng(0)
    rgb = (rng.random((9, 13, 3)) * np.array([1.0, 100.0, 1e4])).astype(np.float32)
    rgb[2, 3] = 0.0
    dec = write_flat_hdr(str(tmp_path / "a.hdr"), rgb)
    got = rt.read_image_float(str(tmp_path / "a.hdr"), flip_y=True)
    assert got.shape == (9, 13, 3) and np.array_equal(got, dec[::-1])               # row 0 = bottom row, as the reference loads it
    assert np.array_equal(rt.read_image_float(str(tmp_path / "a.hdr"), flip_y=False), dec)
    assert np.abs(got - rgb[::-1]).max() <= rgb.max() / 128.0
    (tmp_path / "bad.hdr").write_bytes(b"P6\n1 1\n255\nxxx")
    with pytest.raises(rt.B200RTError):
        rt.read_image_float(str(tmp_path / "bad.hdr"))


def test_hdr_reader_equals_the_reference_loader(rt, tmp_path):
    """The reference's own write_image_hdr -> Utils::read_image_float round trip (stb: run-length encoded scanlines) through
    oracle/_ref against b200rt_hdr_load on the same file: bit for bit."""
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    from sycl_ray_tracing_b200 import scenes
    L = O.RefOracle().lib
    if not hasattr(L, "refo_write_hdr"):
        pytest.skip("oracle/_ref predates the HDR entry points")
    FP = C.POINTER(C.c_float)
    L.refo_write_hdr.argtypes = [FP, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.refo_read_image_float.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, FP]
    for w, h in ((256, 128), (6, 5), (40, 3)):
        env = np.ascontiguousarray(scenes.procedural_sky(256, 128)[:h, :w])
        path = str(tmp_path / f"sky_{w}.hdr").encode()
        assert L.refo_write_hdr(env.ctypes.data_as(FP), w, h, path, 1) == 0
        ref = np.zeros((h, w, 4), np.float32)
        assert L.refo_read_image_float(path, 1, w, h, ref.ctypes.data_as(FP)) == 0
        mine = rt.read_image_float(path.decode())
        assert np.array_equal(mine.view(np.uint32), np.ascontiguousarray(ref[..., :3]).view(np.uint32)), (w, h)
        assert (ref[..., 3] == 0).all()             # read_image_float's alpha (utils.cpp:119); the device expansion writes the same 0
