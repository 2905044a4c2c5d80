"""GPU: entry points added at the end of round 1 (esikf_vio_warp_affine and the shim's per-patch getImagePatch / warpAffine
mirrors). Kept in a file that sorts after the other GPU tests: they reuse kernels that are covered above, but the host
plumbing itself has not run on a GPU yet."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_bind as O
from fast_livo2_b200 import api
from test_gpu_vio import _gpu_warp, _setup, _vio_prior

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_warp_affine_alone_with_caller_matrices(gpu_ctx, small_vio_frame):
    """esikf_vio_warp_affine (include/vio.h:161-162): the same kernel as warp_patches fed with caller-provided matrices —
    identical to the batched producer for its own matrices, and equal to the oracle's warpAffine for arbitrary ones
    (rotation + anisotropic scale, search levels 0 / 1; a zero matrix leaves zeros)."""
    fr = small_vio_frame
    _setup(gpu_ctx, fr)
    prior = _vio_prior(fr)
    g = _gpu_warp(gpu_ctx, fr, prior)
    n = len(fr["vis_pos"])
    again = gpu_ctx.vio_warp_affine(np.zeros(n, np.int32), fr["px_ref"], g["A_cur_ref"], g["search_levels"])
    assert np.array_equal(again, g["warp_patch"])
    m = 24
    rng = np.random.default_rng(5)
    ang = rng.uniform(-0.6, 0.6, m)
    A = np.stack([np.stack([1.3 * np.cos(ang), -0.8 * np.sin(ang)], 1), np.stack([1.3 * np.sin(ang), 0.8 * np.cos(ang)], 1)], 1)  # (m, 2, 2)
    A[-1] = 0.0  # singular with A_ref_cur(0, 0) = NaN: the patch is left untouched (vio.cpp:297-301)
    sl = (np.arange(m) % 2).astype(np.int32)
    out = gpu_ctx.vio_warp_affine(np.zeros(m, np.int32), fr["px_ref"][:m], A, sl)
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    exact = 0
    for i in range(m - 1):
        ref = vio.warp_affine(fr["img_ref"], A[i], fr["px_ref"][i], sl[i])
        np.testing.assert_allclose(out[i], ref, atol=2e-3)
        exact += np.array_equal(out[i], ref)
    assert exact >= m // 2  # float products without contraction on both sides: mostly bit-exact
    assert not out[-1].any()


def test_shim_patch_helpers_match_oracle(small_vio_frame):
    """VIOManager::getImagePatch / VIOManager::warpAffine mirrors (include/vio.h:151, 161-162): one patch each through the
    shim class; only the addressed pyramid level of the caller's buffer is written."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fr = small_vio_frame
    shim = C.CDLL(os.path.join(ROOT, "fast_livo2_b200", "libfl2_shim.so"))
    L = fr["vio_cfg"].levels
    cam = api.CameraC(fr["cam_cfg"].model, fr["cam_cfg"].width, fr["cam_cfg"].height, 0, fr["cam_cfg"].fx, fr["cam_cfg"].fy, fr["cam_cfg"].cx, fr["cam_cfg"].cy)
    cam.d[:] = list(fr["cam_cfg"].d)
    vcfg = api.VioCfgC(fr["vio_cfg"].img_point_cov, L, fr["vio_cfg"].max_iterations, int(fr["vio_cfg"].exposure_estimate_en), 0)
    img, img_ref = np.ascontiguousarray(fr["img"]), np.ascontiguousarray(fr["img_ref"])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    pc = np.array([301.37, 222.81])
    A = np.array([[1.1, -0.2], [0.15, 0.9]])
    px_ref = np.ascontiguousarray(fr["px_ref"][0], dtype=np.float64)
    for level, search_level in ((0, 0), (2, 1), (L - 1, 0)):
        patch = np.full(64 * L, -7.0, np.float32)
        warp = np.full(64 * L, -7.0, np.float32)
        rc = shim.fl2_shim_patch_helpers(C.byref(cam), C.byref(vcfg), vp(img), img.shape[1], img.shape[0], vp(pc), level, vp(patch), vp(A), vp(img_ref), vp(px_ref),
                                         search_level, level, vp(warp))
        assert rc == 0
        sel = slice(64 * level, 64 * level + 64)
        assert np.array_equal(patch[sel], vio.get_image_patch(img, pc, level))
        np.testing.assert_allclose(warp[sel], vio.warp_affine(img_ref, A, px_ref, search_level)[sel], atol=2e-3)
        untouched = np.ones(64 * L, bool)
        untouched[sel] = False
        assert (patch[untouched] == -7.0).all() and (warp[untouched] == -7.0).all()
