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
