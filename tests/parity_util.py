"""Shared comparison helpers for the GPU-vs-oracle parity tests."""
import numpy as np

import oracle_bind as O
from fast_livo2_b200 import synthetic as S

# north_star tolerance: "bit-identical for voxel indexing, within 1e-5 relative on the pose/covariance".
# The tests hold the CUDA path to much tighter bounds (fp64 reduction-order noise only).
POSE_RTOL = 1e-9
COV_RTOL = 1e-6
INFO_RTOL = 1e-10


def pose_diff(a, b):
    ua, ub = S.unpack_state(a), S.unpack_state(b)
    rot = O.rot_err(ua["R"], ub["R"])
    pos = float(np.linalg.norm(ua["p"] - ub["p"]) / max(np.linalg.norm(ub["p"]), 1e-3))
    rest = float(np.abs(a[12:25] - b[12:25]).max())
    cov = float(np.abs(ua["cov"] - ub["cov"]).max() / np.abs(ub["cov"]).max())
    return rot, pos, rest, cov


def assert_state_close(gpu, ref, rot_tol=POSE_RTOL, pos_tol=POSE_RTOL, cov_tol=COV_RTOL, rest_tol=1e-9):
    rot, pos, rest, cov = pose_diff(gpu, ref)
    assert rot < rot_tol, f"rotation differs by {rot} rad"
    assert pos < pos_tol, f"position differs by {pos} (relative)"
    assert rest < rest_tol, f"expo/v/bias/gravity differ by {rest}"
    assert cov < cov_tol, f"covariance differs by {cov} (relative to max)"


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
