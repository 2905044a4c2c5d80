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


@pytest.mark.parametrize("seed,n_pts,n_map,scale", [(4, 20000, 150_000, 0.5), (12, 260_000, 1_000_000, 1.0)])
def test_lio_dealt_schedule_matches_contiguous_and_oracle(gpu_ctx, seed, n_pts, n_map, scale):
    """32-point chunks dealt round-robin over the CTAs: the association is identical, the state agrees to the summation-
    order level with the contiguous schedule and within the usual tolerances with the oracle."""
    import oracle_bind as O
    from fast_livo2_b200 import synthetic as S
    from parity_util import assert_state_close

    kw = dict(lio=S.LioCfg(beam_err=0.01)) if n_pts > 100_000 else dict(scene_scale=scale)
    fr = get_frame(seed=seed, n_pts=n_pts, n_map=n_map, **kw)
    gpu_ctx.set_extrinsics(fr["ext"])
    gpu_ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
    try:
        gpu_ctx.set_tuning(0)
        a = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
        gpu_ctx.set_tuning(api.TUNE_DEAL_POINTS)
        b = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
        b2 = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
    finally:
        gpu_ctx.set_tuning(0)
    assert a["iters"] == b["iters"]
    _bits_equal(a, b, ("match_plane", "normal_plane", "dis_to_plane", "M", "converged"))
    _bits_equal(b, b2, ("state", "HTH", "HTz", "match_plane"))  # deterministic
    assert_state_close(b["state"], a["state"])
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(fr["map"])
    o = lio.state_estimation(fr["pts"], fr["state_prior"], fr["state_prior"])
    assert np.array_equal(b["match_plane"], o["match_plane"])
    assert_state_close(b["state"], o["state"])


def test_bit_identical_tuning_variants(gpu_ctx, small_vio_frame):
    """ESIKF_TUNE_DEFER_DIAGNOSTICS only moves CTA 0's diagnostics writes into the next barrier wait and
    ESIKF_TUNE_VIO_FAST_PATH only caches per-patch inputs / replaces power-of-two divisions / overlaps the boxminus: every
    output of the LIO and VIO updates, diagnostics included, must be bit-identical to the default, in any combination."""
    fr = get_frame(seed=4, n_pts=20000, n_map=150_000, scene_scale=0.5)
    fv = small_vio_frame
    lio_keys = ("state", "match_plane", "normal_plane", "dis_to_plane", "M", "HTH", "HTz", "solution", "total_residual", "converged")
    vio_keys = ("state", "errors", "iters_per_level", "accepted_per_level", "error_trace", "HTH", "HTz", "solution")
    D, F = api.TUNE_DEFER_DIAGNOSTICS, api.TUNE_VIO_FAST_PATH
    out = []
    try:
        for flags in (0, D, F, D | F, D | F):
            gpu_ctx.set_tuning(flags)
            gpu_ctx.set_extrinsics(fr["ext"])
            gpu_ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
            r = gpu_ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
            _setup(gpu_ctx, fv)
            prior = _vio_prior(fv)
            w = _gpu_warp(gpu_ctx, fv, prior)
            args = (fv["img"], fv["vis_pos"], w["warp_patch"], w["search_levels"], fv["inv_ref_expo"], prior, prior)
            v = gpu_ctx.vio_update(*args)
            gpu_ctx.set_solve_mode(1)
            v_lit = gpu_ctx.vio_update(*args)
            gpu_ctx.set_solve_mode(0)
            out.append((r, v, v_lit))
    finally:
        gpu_ctx.set_tuning(0)
        gpu_ctx.set_solve_mode(0)
    for r, v, v_lit in out[1:]:
        assert r["iters"] == out[0][0]["iters"] and v["total_iters"] == out[0][1]["total_iters"]
        _bits_equal(out[0][0], r, lio_keys)
        _bits_equal(out[0][1], v, vio_keys)
        _bits_equal(out[0][2], v_lit, vio_keys)


def test_vio_fast_path_with_two_patches_per_warp_and_search_levels(gpu_ctx):
    """More patches than warps (nothing stays cached) and non-zero search levels / a distorted camera through the FAST path."""
    from fast_livo2_b200 import synthetic as S

    cam = S.CamCfg(width=612, height=512, fx=612.0 * 0.72, fy=612.0 * 0.72, cx=306.0, cy=256.0)
    vcfg = S.VioCfg(levels=5, img_point_cov=1000.0)
    fr = get_frame(seed=13, n_pts=1000, n_map=300_000, n_patches=4000, scene_scale=0.7, cam=cam, vio=vcfg)
    _setup(gpu_ctx, fr)
    prior = _vio_prior(fr, 13)
    w = _gpu_warp(gpu_ctx, fr, prior)
    sl = (np.arange(len(fr["vis_pos"])) % 2).astype(np.int32)  # exercise search_level 1 as well
    args = (fr["img"], fr["vis_pos"], w["warp_patch"], sl, fr["inv_ref_expo"], prior, prior)
    keys = ("state", "errors", "iters_per_level", "accepted_per_level", "error_trace", "HTH", "HTz", "solution")
    try:
        gpu_ctx.set_tuning(0)
        a = gpu_ctx.vio_update(*args)
        gpu_ctx.set_tuning(api.TUNE_VIO_FAST_PATH | api.TUNE_DEFER_DIAGNOSTICS)
        b = gpu_ctx.vio_update(*args)
    finally:
        gpu_ctx.set_tuning(0)
    assert a["total_iters"] == b["total_iters"]
    _bits_equal(a, b, keys)
