import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def rt():
    import sycl_ray_tracing_b200 as m
    return m


@pytest.fixture(scope="session")
def golden_scenes():
    return load_golden("scenes.npz")


@pytest.fixture(scope="session")
def golden_cameras():
    return load_golden("cameras.npz")


def scene_arrays(golden_scenes, key):
    return {k: golden_scenes[f"{key}_{k}"] for k in ("tri9", "mat_idx", "mats10", "emissive")}


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def ensure_built():
    """Build the in-tree native pieces if a test session starts from a clean checkout."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b200rt_build", os.path.join(ROOT, "sycl-ray-tracing_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from oracle import oracle as O
    if not O.have_port() or (os.path.isdir(O.REF_DIR) and not O.have_ref()):
        O.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    ensure_built()


def brute_force_closest(tri, rays):
    """numpy Moller-Trumbore over every triangle (triangle.h:16-60 in float32, strict '<' with ties to the lower index)."""
    import numpy as np
    f32 = np.float32
    tri = np.ascontiguousarray(tri, f32).reshape(-1, 9)
    a = tri[:, 0:3]; e1 = (tri[:, 3:6] - a).astype(f32); e2 = (tri[:, 6:9] - a).astype(f32)
    best_t = np.full(len(rays), -1.0, f32); best_p = np.full(len(rays), -1, np.int32)
    if len(tri) == 0:
        return best_p, best_t
    for r, (o, d) in enumerate(zip(rays[:, :3], rays[:, 3:])):
        h = np.stack([d[1] * e2[:, 2] - d[2] * e2[:, 1], d[2] * e2[:, 0] - d[0] * e2[:, 2], d[0] * e2[:, 1] - d[1] * e2[:, 0]], 1).astype(f32)
        det = ((e1[:, 0] * h[:, 0]).astype(f32) + (e1[:, 1] * h[:, 1]).astype(f32) + (e1[:, 2] * h[:, 2]).astype(f32)).astype(f32)
        ok = ~((det > -1e-7) & (det < 1e-7))
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            f = (f32(1.0) / det).astype(f32)
            s_ = (o - a).astype(f32)
            u = (f * ((s_[:, 0] * h[:, 0]).astype(f32) + (s_[:, 1] * h[:, 1]).astype(f32) + (s_[:, 2] * h[:, 2]).astype(f32)).astype(f32)).astype(f32)
            ok &= ~((u < 0) | (u > 1))
            q = np.stack([s_[:, 1] * e1[:, 2] - s_[:, 2] * e1[:, 1], s_[:, 2] * e1[:, 0] - s_[:, 0] * e1[:, 2], s_[:, 0] * e1[:, 1] - s_[:, 1] * e1[:, 0]], 1).astype(f32)
            v = (f * ((d[0] * q[:, 0]).astype(f32) + (d[1] * q[:, 1]).astype(f32) + (d[2] * q[:, 2]).astype(f32)).astype(f32)).astype(f32)
            ok &= ~((v < 0) | ((u + v).astype(f32) > 1))
            t = (f * ((e2[:, 0] * q[:, 0]).astype(f32) + (e2[:, 1] * q[:, 1]).astype(f32) + (e2[:, 2] * q[:, 2]).astype(f32)).astype(f32)).astype(f32)
        ok &= t > 1e-7
        if ok.any():
            tt = np.where(ok, t, np.inf)
            i = int(np.argmin(tt))            # first (= lowest index) of the minima
            best_t[r] = tt[i]; best_p[r] = i
    return best_p, best_t
