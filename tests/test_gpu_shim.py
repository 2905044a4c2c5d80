"""GPU: the C++ shim (VoxelMapManager::StateEstimation / VIOManager::computeJacobianAndUpdateEKF mirrors) gives the same
states as the direct C-ABI calls and as the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_bind as O
from fast_livo2_b200 import api
from fast_livo2_b200 import synthetic as S
from parity_util import assert_state_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_lio_and_vio_match_oracle(gpu_ctx, small_vio_frame):
    fr = small_vio_frame
    shim = C.CDLL(os.path.join(ROOT, "fast_livo2_b200", "libfl2_shim.so"))
    m = fr["map"]
    k, f, c, p = (np.ascontiguousarray(m["keys"]), np.ascontiguousarray(m["first"]), np.ascontiguousarray(m["count"]), np.ascontiguousarray(m["planes"]))
    lcfg = api.lio_cfg_c(fr["lio_cfg"])
    ext = api.ExtrinsicsC()
    ext.extR[:] = fr["ext"].extR.reshape(9)
    ext.extT[:] = fr["ext"].extT
    ext.Rcl[:] = fr["ext"].Rcl.reshape(9)
    ext.Pcl[:] = fr["ext"].Pcl
    cam = api.CameraC(fr["cam_cfg"].model, fr["cam_cfg"].width, fr["cam_cfg"].height, 0, fr["cam_cfg"].fx, fr["cam_cfg"].fy, fr["cam_cfg"].cx, fr["cam_cfg"].cy)
    cam.d[:] = list(fr["cam_cfg"].d)
    vcfg = api.VioCfgC(fr["vio_cfg"].img_point_cov, fr["vio_cfg"].levels, fr["vio_cfg"].max_iterations, int(fr["vio_cfg"].exposure_estimate_en), 0)
    pts = np.ascontiguousarray(fr["pts"])
    n = len(pts)
    # oracle LIO first: its posterior is what the warp patches are built for
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(m)
    o = lio.state_estimation(pts, fr["state_prior"], fr["state_prior"])
    w = O.oracle_warp_patches(fr, o["state"])
    npatch = len(fr["vis_pos"])
    lio_out, vio_out = np.zeros(386), np.zeros(386)
    neff, nptpl = C.c_int32(0), C.c_int32(0)
    normals = np.zeros((n, 3))
    errs = np.zeros(npatch, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    img = np.ascontiguousarray(fr["img"])
    pos, wp, sl, ie = (np.ascontiguousarray(fr["vis_pos"]), np.ascontiguousarray(w["warp_patch"]), np.ascontiguousarray(w["search_levels"]),
                       np.ascontiguousarray(fr["inv_ref_expo"]))
    sp = np.ascontiguousarray(fr["state_prior"])
    rc = shim.fl2_shim_run(vp(k), vp(f), vp(c), len(f), vp(p), len(p), C.byref(lcfg), C.byref(ext), vp(pts), n, vp(sp), vp(sp), vp(lio_out), C.byref(neff),
                           C.byref(nptpl), vp(normals), C.byref(cam), C.byref(vcfg), vp(img), npatch, vp(pos), vp(wp), vp(sl), vp(ie), vp(vio_out), vp(errs))
    assert rc == 0
    assert neff.value == o["M"][-1] and nptpl.value == o["M"][-1]
    assert_state_close(lio_out, o["state"])
    # pv.normal: the matched plane's normal, zero when the point never matched
    want = np.where(o["normal_plane"][:, None] >= 0, m["planes"]["normal"][np.maximum(o["normal_plane"], 0)], 0.0)
    assert np.array_equal(normals, want)
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    ov = vio.update(img, pos, wp, sl, ie, lio_out, lio_out)
    assert_state_close(vio_out, ov["state"], rot_tol=1e-8, pos_tol=1e-8, cov_tol=1e-6, rest_tol=1e-8)
    np.testing.assert_allclose(errs, ov["errors"], rtol=2e-6, atol=1e-3)


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
