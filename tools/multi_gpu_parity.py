"""Sharded-update parity at N ranks (run under torchrun, one rank per GPU):

  1. every rank first runs the SAME frame on its own GPU un-sharded (a second context without peers): the N = 1 result;
  2. then the sharded update (peer mailboxes or NCCL) in every requested mode; the sharded result must reproduce the
     N = 1 result: per-iteration matched counts and VIO iteration counts identical, per-point association identical on
     the rank's shard, states within 1e-12 (fp64 summation order only), and every rank bit-identical to rank 0;
  3. STEPS free-running LIO + VIO updates (the bench pattern: no host sync between them) must each reproduce the first.

Exit code 1 on any mismatch; one PARITY line per mode on rank 0.

  torchrun --nproc-per-node 4 tools/multi_gpu_parity.py            # config 2 (100 k points + 2 k patches)
  PARITY_CFG=small|cfg2|cfg4|cfg5  PARITY_MODES=p2p,p2p_mode1,nccl  PARITY_STEPS=30
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_livo2_b200 import api, synthetic as S  # noqa: E402
from fast_livo2_b200 import workloads as W  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

cfg_name = os.environ.get("PARITY_CFG", "cfg2")
modes = os.environ.get("PARITY_MODES", "p2p").split(",")
steps = int(os.environ.get("PARITY_STEPS", "30"))
if world > 1 and rank != 0:
    dist.barrier()
fr = W.frame(cfg_name)  # one rank builds (or finds) the seeded frame, the others read the same pickle
if world > 1 and rank == 0:
    dist.barrier()
n, npatch = len(fr["pts"]), len(fr["vis_pos"])
failures = []


def fail(msg):
    failures.append(msg)
    print(f"[rank {rank}] MISMATCH: {msg}", flush=True)


def setup(ctx):
    ctx.set_extrinsics(fr["ext"])
    ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
    if npatch:
        ctx.vio_set_camera(fr["cam_cfg"], fr["vio_cfg"])


def warp(ctx, post_state):
    post = S.unpack_state(post_state)
    ctx.vio_set_image(fr["img"])
    ctx.vio_set_ref_images([fr["img_ref"]])
    T_cur = api.pack_T(*S.camera_pose(fr["ext"], post["R"], post["p"]))
    T_ref = np.tile(api.pack_T(*fr["T_ref"]), (npatch, 1))
    return ctx.vio_warp_patches(np.zeros(npatch, np.int32), fr["px_ref"], fr["vis_pos"], fr["vis_normal"], T_ref, T_cur)


def one_update(ctx, w, post_state):
    ctx.lio_run(fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
    rl = ctx.lio_fetch()
    rv = None
    if npatch:
        ctx.vio_run(post_state, post_state)
        rv = ctx.vio_fetch()
    return rl, rv


# ---- 1. un-sharded result on this GPU
ref_ctx = api.Context(local)
setup(ref_ctx)
ref_ctx.lio_set_scan(fr["pts"])
ref_ctx.lio_run(fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
ref_l = ref_ctx.lio_fetch()
post_state = ref_l["state"].copy()
w = None
ref_v = None
if npatch:
    w = warp(ref_ctx, post_state)
    ref_ctx.vio_set_patches(fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"])
    ref_ctx.vio_run(post_state, post_state)
    ref_v = ref_ctx.vio_fetch()
if rank == 0:
    print(f"N=1 reference: LIO iters {ref_l['iters']} M {ref_l['M'].tolist()}" + (f"  VIO iters {ref_v['total_iters']} per level {ref_v['iters_per_level'][:fr['vio_cfg'].levels].tolist()}" if npatch else ""), flush=True)
ref_ctx.close()


def bcast_equal(arr, what):
    if world == 1:
        return
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    r0 = t.clone()
    dist.broadcast(r0, src=0)
    if not torch.equal(t, r0):
        fail(f"{what}: rank {rank} differs from rank 0")


def compare(tag, rl, rv, strict_first=None):
    beg, cnt = api.shard_range(n, rank, world)
    if rl["iters"] != ref_l["iters"] or rl["M"].tolist() != ref_l["M"].tolist():
        fail(f"{tag}: LIO iters / M {rl['iters']} {rl['M'].tolist()} vs N=1 {ref_l['iters']} {ref_l['M'].tolist()}")
    for key in ("match_plane", "normal_plane", "dis_to_plane"):
        a, b = rl[key][beg:beg + cnt], ref_l[key][beg:beg + cnt]
        if not np.array_equal(a, b):
            bad = np.nonzero(a != b)[0]
            fail(f"{tag}: {key} differs on {len(bad)} points of the shard [{beg},{beg + cnt}), first local indices {bad[:8].tolist()}")
    d = np.abs(rl["state"] - ref_l["state"])
    scale = np.maximum(np.abs(ref_l["state"]), 1e-3)
    if (d[:25] / scale[:25]).max() > 1e-11 or d[25:].max() / np.abs(ref_l["state"][25:]).max() > 1e-11:
        fail(f"{tag}: LIO state differs from N=1 by {(d[:25] / scale[:25]).max():.3e} (pose part) / {d[25:].max():.3e} (cov abs)")
    bcast_equal(rl["state"], f"{tag}: LIO posterior")
    if rv is not None:
        if rv["total_iters"] != ref_v["total_iters"] or rv["iters_per_level"].tolist() != ref_v["iters_per_level"].tolist():
            fail(f"{tag}: VIO iterations {rv['total_iters']} {rv['iters_per_level'].tolist()} vs N=1 {ref_v['total_iters']} {ref_v['iters_per_level'].tolist()}")
        d = np.abs(rv["state"] - ref_v["state"])
        scale = np.maximum(np.abs(ref_v["state"]), 1e-3)
        if (d[:25] / scale[:25]).max() > 1e-10 or d[25:].max() / np.abs(ref_v["state"][25:]).max() > 1e-10:
            fail(f"{tag}: VIO state differs from N=1 by {(d[:25] / scale[:25]).max():.3e}")
        bcast_equal(rv["state"], f"{tag}: VIO posterior")
    if strict_first is not None:
        fl, fv = strict_first
        if not np.array_equal(rl["state"], fl["state"]) or rl["M"].tolist() != fl["M"].tolist():
            fail(f"{tag}: LIO result is not bit-identical to the first update of this mode")
        if rv is not None and (not np.array_equal(rv["state"], fv["state"]) or rv["total_iters"] != fv["total_iters"]):
            fail(f"{tag}: VIO result is not bit-identical to the first update of this mode")


for mode in modes:
    ctx = api.Context(local)
    if world > 1:
        if mode.startswith("p2p"):
            handles = [None] * world
            dist.all_gather_object(handles, ctx.peer_export())
            ctx.peer_attach(rank, world, handles)
        else:
            uid = [api.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(rank, world, uid[0])
    if mode == "p2p_mode1":
        ctx.set_loop_mode(1)
    setup(ctx)
    ctx.lio_set_scan(fr["pts"])
    if npatch:
        ctx.vio_set_image(fr["img"])
        ctx.vio_set_patches(fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"])
    first = one_update(ctx, w, post_state)
    compare(f"{mode} first", *first)
    # free-running updates (no fetch / sync between LIO and VIO: the bench pattern), checked at the end and every 10th
    for k in range(steps):
        ctx.lio_run(fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
        if npatch:
            ctx.vio_run(post_state, post_state)
        if k % 10 == 9 or k == steps - 1:
            rv = ctx.vio_fetch() if npatch else None
            rl = ctx.lio_fetch()
            rl = dict(rl, state=first[0]["state"]) if npatch else rl  # the state buffer is shared: after VIO it holds the VIO posterior
            compare(f"{mode} free-running step {k}", rl, rv, strict_first=first)
    ctx.close()
    ok = torch.tensor([0 if failures else 1], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"PARITY mode={mode} cfg={cfg_name} world={world} steps={steps}: {'OK' if ok.item() else 'FAILED'}  LIO M {first[0]['M'].tolist()}" +
              (f" VIO iters {first[1]['total_iters']}" if npatch else ""), flush=True)

bad = torch.tensor([len(failures)], device="cuda")
if world > 1:
    dist.all_reduce(bad, op=dist.ReduceOp.SUM)
    dist.barrier()
    dist.destroy_process_group()
sys.exit(1 if bad.item() else 0)
