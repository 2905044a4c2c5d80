"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the shard ranges the library hands to each rank tile the
scan exactly, and summing per-shard information buffers with an all-reduce reproduces the unsharded H^T R^-1 H, H^T R^-1 z
and matched count — after which every rank's (redundant) solve sees identical inputs."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_bind as O
from conftest import get_frame
from fast_livo2_b200 import api


def _info_from_rows(sp, lo, hi):
    H, w, z = sp["H"][lo:hi], sp["R_inv"][lo:hi], -sp["dis"][lo:hi].astype(np.float64)
    m = sp["plane"][lo:hi] >= 0
    H, w, z = H[m], w[m], z[m]
    return np.concatenate([(H.T * w) @ H, ((H.T * w) @ z)[:, None]], 1).reshape(-1), int(m.sum())


def _worker(rank, world, port, n, sp, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, c = api.shard_range(n, rank, world)
    info, cnt = _info_from_rows(sp, b, b + c)
    t = torch.from_numpy(np.concatenate([info, [float(cnt)]]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out[rank] = (b, c, t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_information_allreduce_matches_unsharded():
    fr = get_frame(seed=1, n_pts=4000, n_map=150_000, scene_scale=0.5)
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(fr["map"])
    sp = lio.single_pass(fr["pts"], fr["state_prior"], fr["state_prior"])
    n = len(fr["pts"])
    full, cnt = _info_from_rows(sp, 0, n)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, n, sp, out), nprocs=2, join=True)
    (b0, c0, t0), (b1, c1, t1) = out[0], out[1]
    assert b0 == 0 and b1 == c0 and c0 + c1 == n  # contiguous tiling
    assert np.array_equal(t0, t1)                   # every rank holds the same reduced buffer
    assert t0[-1] == cnt
    np.testing.assert_allclose(t0[:-1], full, rtol=1e-12, atol=1e-9)


def test_shard_ranges_tile_any_size():
    for n in (0, 1, 7, 100_000, 260_001):
        for world in (1, 2, 4, 8):
            pos = 0
            for r in range(world):
                b, c = api.shard_range(n, r, world)
                assert b == pos and c in (n // world, n // world + 1)
                pos += c
            assert pos == n
