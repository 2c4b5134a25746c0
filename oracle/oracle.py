"""TEST INFRASTRUCTURE ONLY — ctypes wrappers around the two CPU checkers.

* ``RefOracle``  : oracle/_ref/libref_oracle.so — the UNMODIFIED reference sources behind oracle/ref_driver.cpp.
* ``PortOracle`` : oracle/libpt_oracle.so — the C restatement oracle/pt_oracle.c (pinned bit-exact against RefOracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this module.
The product path (sycl-ray-tracing_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "libpt_oracle.so")
REF_DIR = os.environ.get("REF_DIR", "/root/reference")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(port: bool = True, ref: bool = True) -> None:
    """Compile the checkers (oracle/Makefile). `ref` is a no-op when /root/reference is absent."""
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    if targets:
        subprocess.run(["make", "-s", "-C", HERE, f"REF_DIR={REF_DIR}"] + targets, check=True)


def _opt_f32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _opt_i32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class _OracleBase:
    """Shared ctypes plumbing: both libraries export the same entry points under a different prefix."""

    prefix = ""
    so_path = ""

    def __init__(self):
        if not os.path.exists(self.so_path):
            raise FileNotFoundError(f"{self.so_path} is not built (run oracle.build())")
        self.lib = C.CDLL(self.so_path)
        p = self.prefix
        L = self.lib
        self._fn = {}
        VP, FP, IP, I, F, D = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.c_float, C.c_double

        def reg(name, restype, argtypes):
            fn = getattr(L, p + name)
            fn.restype = restype
            fn.argtypes = argtypes
            self._fn[name] = fn

        reg("max_threads", I, [])
        reg("scene_from_arrays", VP, [FP, I, IP, FP, I, IP, I, FP, IP, I, IP, I])
        reg("scene_free", None, [VP])
        reg("scene_counts", None, [VP, IP])
        reg("set_env", None, [VP, FP, I, I])
        reg("get_env_cdf", None, [VP, FP])
        reg("camera_rays", None, [FP, I, I, FP, I, FP])
        reg("trace", I, [VP, FP, I, I, IP, FP, FP, I])
        reg("primary", D, [VP, FP, I, I, I, IP, FP, I])
        reg("render", D, [VP, FP, I, I, I, I, FP, I])
        reg("render_crop", D, [VP, FP, I, I, I, I, I, I, I, I, FP, I])
        reg("xorshift", C.c_uint32, [C.c_uint32, I, I, FP])

    def max_threads(self) -> int:
        return int(self._fn["max_threads"]())

    # ---- scenes -------------------------------------------------------------------------------------------------
    def scene_from_arrays(self, tri9, mat_idx, mats10, emissive, spheres4=None, sphere_prim=None, sphere_mat_idx=None):
        tri9 = np.ascontiguousarray(tri9, np.float32).reshape(-1, 9)
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        mats10 = np.ascontiguousarray(mats10, np.float32).reshape(-1, 10)
        emissive = np.ascontiguousarray(emissive, np.int32)
        spheres4 = np.zeros((0, 4), np.float32) if spheres4 is None else np.ascontiguousarray(spheres4, np.float32).reshape(-1, 4)
        sphere_prim = np.zeros(0, np.int32) if sphere_prim is None else np.ascontiguousarray(sphere_prim, np.int32)
        sphere_mat_idx = np.zeros(0, np.int32) if sphere_mat_idx is None else np.ascontiguousarray(sphere_mat_idx, np.int32)
        assert len(mat_idx) == len(tri9)
        h = self._fn["scene_from_arrays"](
            _opt_f32(tri9), len(tri9), _opt_i32(mat_idx), _opt_f32(mats10), len(mats10),
            _opt_i32(emissive), len(emissive), _opt_f32(spheres4), _opt_i32(sphere_prim), len(spheres4),
            _opt_i32(sphere_mat_idx), len(sphere_mat_idx))
        if not h:
            raise RuntimeError("scene_from_arrays failed")
        return OracleScene(self, h)


class OracleScene:
    def __init__(self, oracle: _OracleBase, handle):
        self.o = oracle
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                self.o._fn["scene_free"](self.h)
                self.h = None
        except Exception:
            pass

    def counts(self):
        out = np.zeros(7, np.int32)
        self.o._fn["scene_counts"](self.h, _opt_i32(out))
        return dict(zip(["n_tri", "n_mat", "n_emissive", "n_mat_idx", "n_spheres", "env_w", "env_h"], out.tolist()))

    def set_env(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        self.o._fn["set_env"](self.h, _opt_f32(rgba), rgba.shape[1], rgba.shape[0])

    def env_cdf(self):
        c = self.counts()
        out = np.zeros(c["env_w"] * c["env_h"], np.float32)
        self.o._fn["get_env_cdf"](self.h, _opt_f32(out))
        return out

    def trace(self, rays6, mode=0, extra=False, nthreads=0):
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        n = len(rays6)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        ex = np.zeros((n, 8), np.float32) if extra else None
        self.o._fn["trace"](self.h, _opt_f32(rays6), n, mode, _opt_i32(prim), _opt_f32(t), _opt_f32(ex), nthreads)
        return (prim, t, ex) if extra else (prim, t)

    def primary(self, cam17, w, h, mode=0, nthreads=0):
        cam17 = np.ascontiguousarray(cam17, np.float32)
        prim = np.zeros(w * h, np.int32)
        t = np.zeros(w * h, np.float32)
        sec = self.o._fn["primary"](self.h, _opt_f32(cam17), w, h, mode, _opt_i32(prim), _opt_f32(t), nthreads)
        return prim.reshape(h, w), t.reshape(h, w), float(sec)

    def render(self, cam17, w, h, spp, bounces, nthreads=0):
        cam17 = np.ascontiguousarray(cam17, np.float32)
        out = np.zeros((h, w, 4), np.float32)
        sec = self.o._fn["render"](self.h, _opt_f32(cam17), w, h, spp, bounces, _opt_f32(out), nthreads)
        if sec < 0:
            raise RuntimeError("render failed (env map not set, or height < 25)")
        return out, float(sec)

    def render_crop(self, cam17, w, h, spp, bounces, x0, y0, x1, y1, nthreads=0):
        cam17 = np.ascontiguousarray(cam17, np.float32)
        out = np.zeros((y1 - y0, x1 - x0, 4), np.float32)
        sec = self.o._fn["render_crop"](self.h, _opt_f32(cam17), w, h, spp, bounces, x0, y0, x1, y1, _opt_f32(out), nthreads)
        if sec < 0:
            raise RuntimeError("render_crop failed (env map not set)")
        return out, float(sec)


class RefOracle(_OracleBase):
    """The compiled reference itself (kind = "reference")."""

    prefix = "refo_"
    so_path = REF_SO
    kind = "reference"

    def __init__(self):
        super().__init__()
        L = self.lib
        VP, FP, IP, I, F = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.c_float
        L.refo_scene_from_obj.restype = VP
        L.refo_scene_from_obj.argtypes = [C.c_char_p]
        L.refo_scene_get.restype = None
        L.refo_scene_get.argtypes = [VP, FP, IP, FP, IP]
        L.refo_camera_preset.restype = I
        L.refo_camera_preset.argtypes = [C.c_char_p, FP]
        L.refo_camera_make.restype = None
        L.refo_camera_make.argtypes = [F, F, F, F, F, F, FP]
        L.refo_golden_counts.restype = None
        L.refo_golden_counts.argtypes = [IP, IP]
        L.refo_golden.restype = None
        L.refo_golden.argtypes = [FP, FP, FP]
        L.refo_octree_stats.restype = None
        L.refo_octree_stats.argtypes = [VP, C.POINTER(C.c_longlong)]
        L.refo_scene_bvh_seconds.restype = C.c_double
        L.refo_scene_bvh_seconds.argtypes = [VP]

    def scene_from_obj(self, path):
        h = self.lib.refo_scene_from_obj(path.encode())
        if not h:
            raise FileNotFoundError(path)
        return OracleScene(self, h)

    def scene_arrays(self, scene: OracleScene):
        c = scene.counts()
        tri9 = np.zeros((c["n_tri"], 9), np.float32)
        mat_idx = np.zeros(c["n_mat_idx"], np.int32)
        mats10 = np.zeros((c["n_mat"], 10), np.float32)
        emissive = np.zeros(c["n_emissive"], np.int32)
        self.lib.refo_scene_get(scene.h, _opt_f32(tri9), _opt_i32(mat_idx), _opt_f32(mats10), _opt_i32(emissive))
        return dict(tri9=tri9, mat_idx=mat_idx, mats10=mats10, emissive=emissive)

    def camera_preset(self, name):
        out = np.zeros(17, np.float32)
        if self.lib.refo_camera_preset(name.encode(), _opt_f32(out)) != 0:
            raise KeyError(name)
        return out

    def camera_make(self, fov_deg, rot_x_deg=0.0, rot_y_deg=0.0, t=(0.0, 0.0, 0.0)):
        out = np.zeros(17, np.float32)
        self.lib.refo_camera_make(fov_deg, rot_x_deg, rot_y_deg, t[0], t[1], t[2], _opt_f32(out))
        return out

    def camera_rays(self, cam17, w, h, xy):
        cam17 = np.ascontiguousarray(cam17, np.float32)
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros((len(xy), 6), np.float32)
        self._fn["camera_rays"](_opt_f32(cam17), w, h, _opt_f32(xy), len(xy), _opt_f32(out))
        return out

    def golden(self):
        nh, nm = C.c_int(), C.c_int()
        self.lib.refo_golden_counts(C.byref(nh), C.byref(nm))
        hit = np.zeros((nh.value, 6), np.float32)
        pts = np.zeros((nh.value, 3), np.float32)
        miss = np.zeros((nm.value, 6), np.float32)
        self.lib.refo_golden(_opt_f32(hit), _opt_f32(pts), _opt_f32(miss))
        return hit, pts, miss

    def xorshift(self, seed, n_warmup, n):
        out = np.zeros(n, np.float32)
        state = self._fn["xorshift"](seed, n_warmup, n, _opt_f32(out))
        return int(state), out

    def octree_stats(self, scene):
        out = (C.c_longlong * 5)()
        self.lib.refo_octree_stats(scene.h, out)
        return dict(zip(["inner", "leaves", "empty_leaves", "max_leaf", "max_depth"], list(out)))


class PortOracle(_OracleBase):
    """The C restatement (kind = "port")."""

    prefix = "pto_"
    so_path = PORT_SO
    kind = "port"

    def __init__(self):
        super().__init__()
        L = self.lib
        VP, FP, IP, I = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int
        L.pto_render_counted.restype = C.c_double
        L.pto_render_counted.argtypes = [VP, FP, I, I, I, I, FP, C.POINTER(C.c_longlong), I]

    def camera_rays(self, cam17, w, h, xy):
        cam17 = np.ascontiguousarray(cam17, np.float32)
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros((len(xy), 6), np.float32)
        self._fn["camera_rays"](_opt_f32(cam17), w, h, _opt_f32(xy), len(xy), _opt_f32(out))
        return out

    def xorshift(self, seed, n_warmup, n):
        out = np.zeros(n, np.float32)
        state = self._fn["xorshift"](seed, n_warmup, n, _opt_f32(out))
        return int(state), out

    def render_counted(self, scene: OracleScene, cam17, w, h, spp, bounces, nthreads=0):
        """render() plus the number of INTERSECT_SCENE-equivalent queries (render_kernel.cpp:504) it issued."""
        cam17 = np.ascontiguousarray(cam17, np.float32)
        out = np.zeros((h, w, 4), np.float32)
        rays = C.c_longlong(0)
        sec = L = self.lib.pto_render_counted(scene.h, _opt_f32(cam17), w, h, spp, bounces, _opt_f32(out), C.byref(rays), nthreads)
        return out, float(sec), int(rays.value)


def quantise_rgba8(image, flip_y: bool = True):
    """numpy restatement of write_image_png's quantisation loop (source/image_io.cpp:165-182 with clamp() :157-162):
    pixel = image * 255 in float, clamp to [0, 255], truncate to unsigned char; flipY writes row 0 last. Pinned against a
    PNG written by the reference itself (tests/golden/quantise.npz). Test infrastructure, like the rest of oracle/."""
    p = np.ascontiguousarray(image, np.float32) * np.float32(255)
    c = np.where(p < 0, np.float32(0), np.where(p > 255, np.float32(255), p)).astype(np.uint8)
    return c[::-1].copy() if flip_y else c


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_port() -> bool:
    return os.path.exists(PORT_SO)


def best_oracle():
    """The compiled reference when its .so travelled with the repo, else the port."""
    return RefOracle() if have_ref() else PortOracle()
