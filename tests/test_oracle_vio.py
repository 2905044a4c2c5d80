"""Oracle VIO checks on CPU: patch extraction / warp against independent numpy formulas, camera models, and the
coarse-to-fine photometric update behaving like an ESIKF update should."""
import numpy as np

import oracle_bind as O
from fast_livo2_b200 import synthetic as S


def _np_bilinear_patch(img, pc, level):
    """Independent numpy restatement of getImagePatch (vio.cpp:203-225) in float32."""
    f32 = np.float32
    scale = 1 << level
    u_ref, v_ref = f32(pc[0]), f32(pc[1])
    u_i = int(np.floor(f32(pc[0] / scale)) * scale)
    v_i = int(np.floor(f32(pc[1] / scale)) * scale)
    su = f32((u_ref - f32(u_i)) / f32(scale))
    sv = f32((v_ref - f32(v_i)) / f32(scale))
    wtl = f32((1.0 - float(su)) * (1.0 - float(sv)))
    wtr = f32(float(su) * (1.0 - float(sv)))
    wbl = f32((1.0 - float(su)) * float(sv))
    wbr = f32(su * sv)
    out = np.zeros(64, f32)
    for x in range(8):
        for y in range(8):
            r, c = v_i - 4 * scale + x * scale, u_i - 4 * scale + y * scale
            a, b, cc, d = (f32(img[r, c]), f32(img[r, c + scale]), f32(img[r + scale, c]), f32(img[r + scale, c + scale]))
            out[x * 8 + y] = f32(f32(f32(wtl * a) + f32(wtr * b)) + f32(wbl * cc)) + f32(wbr * d)
    return out


def test_get_image_patch_matches_numpy(small_vio_frame):
    fr = small_vio_frame
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    rng = np.random.default_rng(0)
    for _ in range(20):
        pc = np.array([rng.uniform(100, 540), rng.uniform(100, 410)])
        for level in (0, 1, 2, 3):
            np.testing.assert_array_equal(vio.get_image_patch(fr["img"], pc, level), _np_bilinear_patch(fr["img"], pc, level))


def test_camera_models_roundtrip():
    ext, vcfg = S.avia_extrinsics(), S.VioCfg()
    cams = [S.CamCfg(), S.CamCfg(d=(-0.076160, 0.123001, -0.00113, 0.000251, 0.0)),
            S.CamCfg(model=1, width=720, height=540, fx=351.314, fy=351.491, cx=367.852, cy=253.840,
                     d=(-0.03696737352869157, -0.008917880497032812, 0.008912969593422046, -0.0037685977496087313, 0.0))]
    rng = np.random.default_rng(1)
    for cam in cams:
        vio = O.OracleVIO(cam, ext, vcfg)
        for _ in range(50):
            px = np.array([rng.uniform(60, cam.width - 60), rng.uniform(60, cam.height - 60)])
            f = vio.cam2world(px)
            assert abs(np.linalg.norm(f) - 1) < 1e-12
            back = vio.world2cam(f * rng.uniform(0.5, 20))
            tol = 1e-9 if cam.d[0] == 0 else (2e-3 if cam.model == 0 else 1e-6)  # cv::undistortPoints: 5 fixed-point iterations
            np.testing.assert_allclose(back, px, atol=tol)
            # generator-side camera (independent) agrees on the projection
            np.testing.assert_allclose(S.cam_project(cam, f[None])[0], vio.world2cam(f), atol=1e-9)


def test_warp_of_identical_views_reproduces_the_patch(small_vio_frame):
    """With T_cur == T_ref the affine warp is the identity and warpAffine equals plain bilinear sampling."""
    fr = small_vio_frame
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    for i in range(0, 40, 3):
        A, sl = vio.warp_matrix(fr["px_ref"][i], fr["vis_pos"][i], fr["vis_normal"][i], fr["T_ref"], fr["T_ref"])
        np.testing.assert_allclose(A, np.eye(2), atol=1e-5)
        assert sl == 0
        wp = vio.warp_affine(fr["img_ref"], np.eye(2), fr["px_ref"][i], 0)
        px = fr["px_ref"][i].astype(np.float32)
        for lvl in range(fr["vio_cfg"].levels):
            for y in range(8):
                for x in range(8):
                    u = np.float32((x - 4) * (1 << lvl)) + px[0]
                    v = np.float32((y - 4) * (1 << lvl)) + px[1]
                    xi, yi = int(np.floor(u)), int(np.floor(v))
                    sx, sy = np.float32(u - xi), np.float32(v - yi)
                    im = fr["img_ref"].astype(np.float32)
                    ref = ((1 - sx) * (1 - sy) * im[yi, xi] + (1 - sx) * sy * im[yi + 1, xi] + sx * (1 - sy) * im[yi, xi + 1] + sx * sy * im[yi + 1, xi + 1])
                    assert abs(wp[64 * lvl + y * 8 + x] - ref) < 1e-3


def test_vio_update_reduces_photometric_error_and_pose_error(small_vio_frame):
    fr = small_vio_frame
    w = O.oracle_warp_patches(fr, fr["state_prior"])
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    # prior: small perturbation of truth (VIO runs after LIO in the reference: LIVMapper.cpp:267-279)
    rng = np.random.default_rng(3)
    t = S.unpack_state(fr["state_true"])
    cov = S.random_prior_cov(rng, scale=0.2)
    prior = S.pack_state(t["R"] @ S.so3_exp(rng.normal(0, np.deg2rad(0.15), 3)), t["p"] + rng.normal(0, 0.01, 3), 1.0, t["v"], g=t["g"], cov=cov)
    w = O.oracle_warp_patches(fr, prior)
    r = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], prior, prior)
    L = fr["vio_cfg"].levels
    assert r["total_iters"] >= L and r["total_iters"] <= L * fr["vio_cfg"].max_iterations
    assert (r["iters_per_level"][:L] >= 1).all() and (r["accepted_per_level"][:L] >= 1).all()  # iteration 0 is always accepted
    e0 = (O.rot_err(t["R"], S.unpack_state(prior)["R"]), np.linalg.norm(t["p"] - S.unpack_state(prior)["p"]))
    post = S.unpack_state(r["state"])
    e1 = (O.rot_err(t["R"], post["R"]), np.linalg.norm(t["p"] - post["p"]))
    assert e1[0] < e0[0] and e1[1] < e0[1]
    # finest level error below coarsest level's first error; covariance shrinks and stays symmetric PD
    assert r["error_trace"][0][r["iters_per_level"][0] - 1] < r["error_trace"][L - 1][0] * 4
    P0, P1 = S.unpack_state(prior)["cov"], post["cov"]
    assert np.linalg.eigvalsh(0.5 * (P1 + P1.T)).min() > 0
    assert (np.diag(P1)[:6] < np.diag(P0)[:6]).all()


def test_vio_rollback_semantics(small_vio_frame):
    """error-gated accept / rollback (vio.cpp:1648-1681): a rejected iteration restores old_state and ends the level."""
    fr = small_vio_frame
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    w = O.oracle_warp_patches(fr, fr["state_true"])
    r = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], fr["state_true"], fr["state_true"])
    for lvl in range(fr["vio_cfg"].levels):
        it, acc = r["iters_per_level"][lvl], r["accepted_per_level"][lvl]
        assert acc in (it, it - 1)
        tr = r["error_trace"][lvl][:it]
        assert all(tr[k + 1] <= tr[k] for k in range(acc - 1))  # accepted errors are non-increasing
        if acc == it - 1:
            assert tr[it - 1] > tr[it - 2]


def test_vio_jacobian_rows_match_finite_differences(small_vio_frame):
    """Finite-difference pin of the geometric Jacobian chain (vio.cpp:1574-1629: Jdpi, [pf]x, Jdphi_dR, Jdp_dR, Jdp_dt).
    On an integer-slope ramp image the bilinear sample and the central-difference gradient are exact, so every one of
    the 64 rows of a patch is d(inv_expo * I(pi(Rcw p + Pcw)))/d(dtheta, dp) and H^T H[:6,:6] = 64 j j^T."""
    import dataclasses

    fr = small_vio_frame
    ext, cam = fr["ext"], fr["cam_cfg"]
    assert cam.model == 0 and not any(cam.d)
    vcfg = dataclasses.replace(fr["vio_cfg"], levels=1, max_iterations=1)
    st = S.unpack_state(fr["state_prior"])
    pos = fr["vis_pos"][3]

    def project(state):
        s = S.unpack_state(state)
        Rcw, Pcw = S.camera_pose(ext, s["R"], s["p"])
        pf = Rcw @ pos + Pcw
        return np.array([cam.fx * pf[0] / pf[2] + cam.cx, cam.fy * pf[1] / pf[2] + cam.cy])

    pc0 = project(fr["state_prior"])
    uc, vc = int(round(pc0[0])), int(round(pc0[1]))
    a, b = 2, -1  # integer slopes: the u8 image is an exact plane around the patch
    uu, vv = np.meshgrid(np.arange(cam.width), np.arange(cam.height))
    img = np.clip(127 + a * (uu - uc) + b * (vv - vc), 0, 255).astype(np.uint8)

    vio = O.OracleVIO(cam, ext, vcfg)
    o = vio.update(img, pos[None], np.zeros((1, 64), np.float32), np.zeros(1, np.int32), np.ones(1), fr["state_prior"], fr["state_prior"])
    HTH = o["HTH"][0][0]

    lib = O.load()

    def boxplus(state, d):
        out = np.zeros_like(state)
        lib.orc_boxplus(O.dptr(O.c64(state)), O.dptr(O.c64(d)), O.dptr(out))
        return out

    h = 1e-4  # above Exp()'s 1e-5 identity threshold (so3_math.h:44-66)
    j = np.zeros(6)
    for k in range(6):
        d = np.zeros(19)
        d[k] = h
        pp, pm = project(boxplus(fr["state_prior"], d)), project(boxplus(fr["state_prior"], -d))
        j[k] = st["inv_expo"] * (a * (pp[0] - pm[0]) + b * (pp[1] - pm[1])) / (2 * h)
    # float32 bilinear weights carry ~1e-7 relative error into du, dv
    np.testing.assert_allclose(HTH[:6, :6], 64.0 * np.outer(j, j), rtol=2e-5, atol=1e-6 * np.abs(j).max() ** 2 * 64)
    # 7th column: the sampled intensity itself (vio.cpp:1626); sum over the ramp patch = 64*127 + offsets
    cur_sum = HTH[:6, 6] / j
    assert np.allclose(cur_sum, cur_sum[0], rtol=1e-4)


def test_inverse_compositional_variant_converges_like_the_forward_one(small_vio_frame):
    """vio/inverse_composition_en (vio.cpp:792-795): gradients of the reference patch instead of the current image. On the
    synthetic frame it must pull the prior to the truth about as well as the forward variant, with a 6-column H."""
    fr = small_vio_frame
    rng = np.random.default_rng(3)
    t = S.unpack_state(fr["state_true"])
    prior = S.pack_state(t["R"] @ S.so3_exp(rng.normal(0, np.deg2rad(0.15), 3)), t["p"] + rng.normal(0, 0.01, 3), 1.0, t["v"], g=t["g"],
                         cov=S.random_prior_cov(rng, scale=0.2))
    w = O.oracle_warp_patches(fr, prior)
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    vio.set_inverse_refs(**O.inverse_refs_from_frame(fr))
    n = len(fr["vis_pos"])
    args = (fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], np.ones(n), prior, prior)
    vio.set_inverse(True)
    inv = vio.update(*args)
    vio.set_inverse(False)
    fwd = vio.update(*args)

    def err(s):
        a = S.unpack_state(s)
        return O.rot_err(a["R"], t["R"]), np.linalg.norm(a["p"] - t["p"])

    e0, ei, ef = err(prior), err(inv["state"]), err(fwd["state"])
    assert inv["total_iters"] >= fr["vio_cfg"].levels and inv["accepted_per_level"][: fr["vio_cfg"].levels].min() >= 1
    assert ei[0] < 0.2 * e0[0] and ei[1] < 0.5 * e0[1]
    assert ei[0] < 5 * ef[0] + 1e-4 and ei[1] < 5 * ef[1] + 1e-3
    Pi = S.unpack_state(inv["state"])["cov"]
    assert np.linalg.eigvalsh(0.5 * (Pi + Pi.T)).min() > 0
    top = fr["vio_cfg"].levels - 1
    assert not inv["HTH"][top][0][6].any() and S.unpack_state(inv["state"])["inv_expo"] != 0
