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
