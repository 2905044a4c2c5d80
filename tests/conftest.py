import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    # a fresh checkout has no built libraries (they are git-ignored): build them once instead of failing every import
    pkg = os.path.join(ROOT, "fast_livo2_b200")
    if not (os.path.exists(os.path.join(pkg, "libesikf_b200.so")) and os.path.exists(os.path.join(pkg, "libfl2_shim.so"))):
        import __graft_entry__

        __graft_entry__.build()


_frames = {}


def get_frame(**kw):
    """Seeded synthetic frames are expensive to build; cache them per parameter set for the session."""
    from fast_livo2_b200 import synthetic as S

    key = tuple(sorted((k, repr(v)) for k, v in kw.items()))
    if key not in _frames:
        _frames[key] = S.cached_frame(**kw)
    return _frames[key]


@pytest.fixture(scope="session")
def small_frame():
    return get_frame(seed=1, n_pts=4000, n_map=150_000, scene_scale=0.5)


@pytest.fixture(scope="session")
def small_vio_frame():
    return get_frame(seed=2, n_pts=2000, n_map=120_000, n_patches=150, scene_scale=0.5)


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fast_livo2_b200 import api

    ctx = api.Context(0)  # raises if libesikf_b200.so is missing: the CUDA path must be the one that runs
    yield ctx
    ctx.close()
