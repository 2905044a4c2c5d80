"""GPU: the replicated-solve persistent kernels (loop_mode 2, the default: every CTA sums the partials and runs the gain
solve itself, one grid barrier per iteration) must reproduce loop_mode 1 (solve on CTA 0) bit for bit — states,
associations and per-iteration diagnostics — in both solve modes, across repeated launches (barrier counters and partial
buffers alternate)."""
import numpy as np
import pytest

from conftest import get_frame
from fast_livo2_b200 import api
from test_gpu_vio import _gpu_warp, _setup, _vio_prior

pytestmark = pytest.mark.gpu


def _bits_equal(a, b, keys):
    for k in keys:
        x, y = np.ascontiguousarray(a[k]), np.ascontiguousarray(b[k])
        assert x.shape == y.shape and x.tobytes() == y.tobytes(), k


@pytest.mark.parametrize("solve_mode", [0, 1])
def test_lio_replicated_solve_is_bit_identical(gpu_ctx, solve_mode):
    fr = get_frame(seed=4, n_pts=20000, n_map=150_000, scene_scale=0.5)
    gpu_ctx.set_extrinsics(fr["ext"])
    gpu_ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
    gpu_ctx.set_solve_mode(solve_mode)
    try:
        out = {}
        for mode in (1, 2, 2, 1, 2):
            gpu_ctx.set_loop_mode(mode)
            r = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
            out.setdefault(mode, []).append(r)
    finally:
        gpu_ctx.set_loop_mode(api.DEFAULT_LOOP_MODE)
        gpu_ctx.set_solve_mode(0)
    ref = out[1][0]
    assert ref["iters"] >= 3
    keys = ("state", "match_plane", "normal_plane", "dis_to_plane", "M", "HTH", "HTz", "solution", "total_residual", "converged")
    for r in out[2] + out[1][1:]:
        assert r["iters"] == ref["iters"]
        _bits_equal(ref, r, keys)


def test_lio_replicated_solve_several_tiles_per_cta(gpu_ctx):
    """260 k points: every CTA walks several tiles, nothing stays resident in the lanes' slots."""
    from fast_livo2_b200 import synthetic as S

    fr = get_frame(seed=12, n_pts=260_000, n_map=1_000_000, lio=S.LioCfg(beam_err=0.01))
    gpu_ctx.set_extrinsics(fr["ext"])
    gpu_ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
    try:
        gpu_ctx.set_loop_mode(1)
        a = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
        gpu_ctx.set_loop_mode(2)
        b = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
    finally:
        gpu_ctx.set_loop_mode(api.DEFAULT_LOOP_MODE)
    assert a["iters"] == b["iters"]
    _bits_equal(a, b, ("state", "match_plane", "normal_plane", "dis_to_plane", "M", "HTH"))


@pytest.mark.parametrize("solve_mode", [0, 1])
def test_vio_replicated_solve_is_bit_identical(gpu_ctx, small_vio_frame, solve_mode):
    fr = small_vio_frame
    _setup(gpu_ctx, fr)
    prior = _vio_prior(fr)
    w = _gpu_warp(gpu_ctx, fr, prior)
    args = (fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], prior, prior)
    gpu_ctx.set_solve_mode(solve_mode)
    try:
        out = {}
        for mode in (1, 2, 2, 1, 2):
            gpu_ctx.set_loop_mode(mode)
            out.setdefault(mode, []).append(gpu_ctx.vio_update(*args))
    finally:
        gpu_ctx.set_loop_mode(api.DEFAULT_LOOP_MODE)
        gpu_ctx.set_solve_mode(0)
    ref = out[1][0]
    assert ref["total_iters"] >= 4
    keys = ("state", "errors", "iters_per_level", "accepted_per_level", "error_trace", "HTH", "HTz", "solution")
    for r in out[2] + out[1][1:]:
        assert r["total_iters"] == ref["total_iters"]
        _bits_equal(ref, r, keys)
