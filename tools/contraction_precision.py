"""Which arithmetic can carry the H^T R^-1 H contraction? CPU experiment on the BASELINE config-2 frame (oracle rows of the
first LIO iteration): the 6 x 6 information matrix, H^T R^-1 z and the resulting first ESIKF solution / posterior covariance
with the contraction done in fp64 (what the kernels do on the DMMA path), fp32, TF32 (10-bit mantissa products, fp32
accumulation) and 3 x TF32 split products — against the north star's 1e-5 budget on pose / covariance.
    python tools/contraction_precision.py        (CPU only, ~1 min)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_bind as O  # noqa: E402
from fast_livo2_b200 import synthetic as S  # noqa: E402


def tf32(x):
    """Round fp32 to TF32 (10 explicit mantissa bits, round to nearest even on the dropped 13 bits)."""
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def contract(A, w, z, kind):
    """sum_i w_i a_i a_i^T and sum_i w_i a_i z_i with products / accumulation in the given arithmetic."""
    if kind == "fp64":
        B = A * w[:, None]
        return B.T @ A, B.T @ z
    A32, w32, z32 = A.astype(np.float32), w.astype(np.float32), z.astype(np.float32)
    B32 = (A32 * w32[:, None]).astype(np.float32)
    if kind == "fp32":
        return (B32.T @ A32).astype(np.float64), (B32.T @ z32).astype(np.float64)
    if kind == "tf32":
        return (tf32(B32).T @ tf32(A32)).astype(np.float64), (tf32(B32).T @ tf32(z32)).astype(np.float64)
    if kind == "3xtf32":  # a = a_hi + a_lo: hi*hi + hi*lo + lo*hi, fp32 accumulation
        Bh, Ah, zh = tf32(B32), tf32(A32), tf32(z32)
        Bl, Al, zl = tf32(B32 - Bh), tf32(A32 - Ah), tf32(z32 - zh)
        HTH = Bh.T @ Ah + Bh.T @ Al + Bl.T @ Ah
        HTz = Bh.T @ zh + Bh.T @ zl + Bl.T @ zh
        return HTH.astype(np.float64), HTz.astype(np.float64)
    raise ValueError(kind)


def solve(HTH, HTz, P):
    H = np.zeros((19, 19))
    H[:6, :6] = HTH
    K1 = np.linalg.inv(H + np.linalg.inv(P))
    G = K1[:, :6] @ HTH
    return K1[:, :6] @ HTz, (np.eye(19) - np.pad(G, ((0, 0), (0, 13)))) @ P


def main():
    fr = S.cached_frame(seed=0, n_pts=100000, n_map=1000000, n_patches=2000)
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(fr["map"])
    sp = lio.single_pass(fr["pts"], fr["state_prior"], fr["state_prior"])
    m = sp["plane"] >= 0
    A, w, z = sp["H"][m], sp["R_inv"][m], -sp["dis"][m].astype(np.float64)
    P = S.unpack_state(fr["state_prior"])["cov"]
    ref = contract(A, w, z, "fp64")
    sol_ref, cov_ref = solve(*ref, P)
    print(f"{m.sum()} matched points; |solution| rot {np.linalg.norm(sol_ref[:3]):.3e} rad, pos {np.linalg.norm(sol_ref[3:6]):.3e} m")
    print(f"{'arithmetic':10s} {'rel err HTH':>12s} {'rel err HTz':>12s} {'rot err [rad]':>14s} {'pos err rel':>12s} {'cov err / max':>14s}")
    for kind in ("fp64", "fp32", "tf32", "3xtf32"):
        HTH, HTz = contract(A, w, z, kind)
        sol, cov = solve(HTH, HTz, P)
        p_ref = S.unpack_state(fr["state_prior"])["p"] + sol_ref[3:6]
        print(f"{kind:10s} {np.abs(HTH - ref[0]).max() / np.abs(ref[0]).max():12.2e} {np.abs(HTz - ref[1]).max() / np.abs(ref[1]).max():12.2e} "
              f"{np.linalg.norm(sol[:3] - sol_ref[:3]):14.2e} {np.linalg.norm(sol[3:6] - sol_ref[3:6]) / np.linalg.norm(p_ref):12.2e} "
              f"{np.abs(cov - cov_ref).max() / np.abs(cov_ref).max():14.2e}")


if __name__ == "__main__":
    main()
