"""Oracle self-checks (CPU): state algebra and dense helpers against independent numpy / scipy formulas."""
import ctypes as C

import numpy as np
from scipy.spatial.transform import Rotation

import oracle_bind as O
from fast_livo2_b200 import synthetic as S


def test_exp_matches_scipy_and_small_angle_quirk():
    lib = O.load()
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = rng.normal(0, 0.7, 3)
        R = np.zeros(9)
        lib.orc_exp(O.dptr(v), O.dptr(R))
        np.testing.assert_allclose(R.reshape(3, 3), Rotation.from_rotvec(v).as_matrix(), atol=1e-13)
    # include/utils/so3_math.h:48 — identity when |v| <= 1e-5
    v = np.array([6e-6, 0, 0.0])
    R = np.zeros(9)
    lib.orc_exp(O.dptr(v), O.dptr(R))
    assert np.array_equal(R.reshape(3, 3), np.eye(3))


def test_log_branches():
    lib = O.load()
    for ang in (0.5, 2e-4, 1e-9):
        Rm = Rotation.from_rotvec([0, 0, ang]).as_matrix().copy()
        v = np.zeros(3)
        lib.orc_log(O.dptr(np.ascontiguousarray(Rm.reshape(9))), O.dptr(v))
        if ang == 1e-9:  # trace > 3 - 1e-6 -> theta = 0 -> 0.5 * K
            np.testing.assert_allclose(v, [0, 0, 0.5 * (Rm[1, 0] - Rm[0, 1])], atol=0)
        else:
            np.testing.assert_allclose(v, [0, 0, ang], rtol=1e-6)


def test_boxplus_boxminus_roundtrip():
    lib = O.load()
    rng = np.random.default_rng(1)
    s = S.pack_state(S.so3_exp(rng.normal(0, 0.3, 3)), rng.normal(size=3), 1.1, rng.normal(size=3), rng.normal(size=3), rng.normal(size=3),
                     rng.normal(size=3), S.random_prior_cov(rng))
    d = rng.normal(0, 0.05, 19)
    out = np.zeros(386)
    lib.orc_boxplus(O.dptr(s), O.dptr(d), O.dptr(out))
    back = np.zeros(19)
    lib.orc_boxminus(O.dptr(out), O.dptr(s), O.dptr(back))
    np.testing.assert_allclose(back, d, atol=1e-12)
    # covariance is carried through untouched
    assert np.array_equal(out[25:], s[25:])


def test_inverse19_matches_numpy():
    lib = O.load()
    rng = np.random.default_rng(2)
    P = S.random_prior_cov(rng)
    inv = np.zeros((19, 19))
    lib.orc_inverse19(O.dptr(np.ascontiguousarray(P)), O.dptr(inv))
    np.testing.assert_allclose(inv @ P, np.eye(19), atol=1e-9)
    np.testing.assert_allclose(inv, np.linalg.inv(P), rtol=1e-8, atol=1e-6)


def test_calc_body_cov_matches_vectorised_generator_restatement():
    lib = O.load()
    rng = np.random.default_rng(3)
    pts = rng.normal(0, 8, (200, 3))
    ref = S.calc_body_cov_np(pts, 0.02, 0.05)
    for i in range(len(pts)):
        cov = np.zeros(9)
        lib.orc_calc_body_cov(O.dptr(np.ascontiguousarray(pts[i])), C.c_float(0.02), C.c_float(0.05), O.dptr(cov), None)
        np.testing.assert_allclose(cov.reshape(3, 3), ref[i], rtol=1e-9, atol=1e-14)
    # symmetric PSD with one range direction and two bearing directions
    w = np.linalg.eigvalsh(ref)
    assert (w > 0).all()


def test_default_state_cov():
    lib = O.load()
    s = np.zeros(386)
    lib.orc_default_state(O.dptr(s))
    cov = s[25:].reshape(19, 19)
    assert cov[0, 0] == 0.01 and cov[6, 6] == 0.00001 and cov[10, 10] == 0.00001 and cov[7, 7] == 0.01
