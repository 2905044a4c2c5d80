"""Gain K_1[:, :m] = ((H^T H padded) + P^-1)^-1 [:, :m]: the reference's literal form (two 19 x 19 inversions in double)
against the push-through form the kernels use, P[:, :m] (I + A P_mm)^-1 (one m x m solve), both measured against an
80-bit long-double evaluation of the literal formula. Inputs: the first LIO iteration of BASELINE config 2 (oracle) and a
VIO-like scaling (P / img_point_cov, m = 7).      python tools/gain_precision.py      (CPU only)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bind as O  # noqa: E402
from fast_livo2_b200 import synthetic as S  # noqa: E402

LD = np.longdouble


def inv_gj(A):
    """Gauss-Jordan with partial pivoting in the dtype of A."""
    n = len(A)
    W = np.concatenate([A.copy(), np.eye(n, dtype=A.dtype)], 1)
    for k in range(n):
        p = k + int(np.argmax(np.abs(W[k:, k])))
        if p != k:
            W[[k, p]] = W[[p, k]]
        W[k] = W[k] / W[k, k]
        for r in range(n):
            if r != k:
                W[r] = W[r] - W[r, k] * W[k]
    return W[:, n:]


def gains(A, P, m):
    Hp = np.zeros((19, 19), A.dtype)
    Hp[:m, :m] = A
    literal = inv_gj(Hp + inv_gj(P))[:, :m]
    push = P[:, :m] @ inv_gj(np.eye(m, dtype=A.dtype) + A @ P[:m, :m])
    return literal, push


def report(name, A, P, m):
    ref, ref_push = gains(A.astype(LD), P.astype(LD), m)
    lit, push = gains(A.astype(np.float64), P.astype(np.float64), m)
    scale = float(np.abs(ref).max())
    print(f"{name:34s} cond(P) {np.linalg.cond(P):9.2e}  literal(double) err {float(np.abs(lit - ref).max()) / scale:9.2e}  "
          f"push-through(double) err {float(np.abs(push - ref).max()) / scale:9.2e}  forms agree in long double to {float(np.abs(ref_push - ref).max()) / scale:9.2e}")


def main():
    fr = S.cached_frame(seed=0, n_pts=100000, n_map=1000000, n_patches=2000)
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(fr["map"])
    r = lio.state_estimation(fr["pts"], fr["state_prior"], fr["state_prior"])
    P = S.unpack_state(fr["state_prior"])["cov"]
    print("relative error of K_1[:, :m] against the long-double literal formula")
    report("LIO iteration 0 (m = 6)", r["HTH"][0], P, 6)
    report("LIO last iteration (m = 6)", r["HTH"][-1], P, 6)
    w = O.oracle_warp_patches(fr, r["state"])
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    v = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], r["state"], r["state"])
    top = fr["vio_cfg"].levels - 1
    Pv = S.unpack_state(r["state"])["cov"] / fr["vio_cfg"].img_point_cov
    report("VIO coarsest level, it 0 (m = 7)", v["HTH"][top][0], Pv, 7)
    report("VIO finest level, it 0 (m = 7)", v["HTH"][0][0], Pv, 7)


if __name__ == "__main__":
    main()
