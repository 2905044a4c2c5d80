"""Run under torchrun (one rank per GPU): sharded LIO + VIO update with the per-iteration all-reduce of the information
buffer; checks every rank ends with the same state and that it matches the single-process CPU oracle."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fast_livo2_b200 import api, synthetic as S
import oracle_bind as O
from parity_util import assert_state_close

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fr = S.cached_frame(seed=4, n_pts=20000, n_map=150_000, n_patches=0, scene_scale=0.5)
frv = S.cached_frame(seed=2, n_pts=2000, n_map=120_000, n_patches=150, scene_scale=0.5)
ctx = api.Context(local)
mode = os.environ.get("ESIKF_COMM", "p2p")
if mode == "p2p":
    handles = [None] * world
    dist.all_gather_object(handles, ctx.peer_export())
    ctx.peer_attach(rank, world, handles)
else:
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
tuning = int(os.environ.get("ESIKF_TUNING", "0"))  # e.g. 8 = TUNE_PEER_REPLICATED (replicated-solve kernels pulling the peer sum)
ctx.set_tuning(tuning)
ctx.set_extrinsics(fr["ext"]); ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size)
g = ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
# every rank holds the same posterior (redundant solve on identical all-reduced information)
t = torch.from_numpy(g["state"]).cuda(); ref = t.clone(); dist.broadcast(ref, src=0)
assert torch.equal(t, ref), "ranks disagree on the LIO posterior"
# per-point outputs: each rank fills its shard; combine with max (unwritten entries are -1 / untouched)
n = len(fr["pts"]); base, rem = n // world, n % world
beg = rank * base + min(rank, rem); cnt = base + (1 if rank < rem else 0)
mp = torch.full((n,), -2, dtype=torch.int32, device="cuda"); mp[beg:beg + cnt] = torch.from_numpy(g["match_plane"][beg:beg + cnt]).cuda()
dist.all_reduce(mp, op=dist.ReduceOp.MAX)
# VIO
ctx.vio_set_camera(frv["cam_cfg"], frv["vio_cfg"]); ctx.vio_set_image(frv["img"]); ctx.vio_set_ref_images([frv["img_ref"]])
st = S.unpack_state(frv["state_true"])
nv = len(frv["vis_pos"])
w = ctx.vio_warp_patches(np.zeros(nv, np.int32), frv["px_ref"], frv["vis_pos"], frv["vis_normal"], np.tile(api.pack_T(*frv["T_ref"]), (nv, 1)),
                         api.pack_T(*S.camera_pose(frv["ext"], st["R"], st["p"])))
ctx.set_extrinsics(frv["ext"])
gv = ctx.vio_update(frv["img"], frv["vis_pos"], w["warp_patch"], w["search_levels"], frv["inv_ref_expo"], frv["state_prior"], frv["state_prior"])
tv = torch.from_numpy(gv["state"]).cuda(); refv = tv.clone(); dist.broadcast(refv, src=0)
assert torch.equal(tv, refv), "ranks disagree on the VIO posterior"
if rank == 0:
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"]); lio.set_map(fr["map"])
    o = lio.state_estimation(fr["pts"], fr["state_prior"], fr["state_prior"])
    assert g["iters"] == o["iters"] and np.array_equal(g["M"], o["M"]), (g["M"], o["M"])
    assert np.array_equal(mp.cpu().numpy(), o["match_plane"])
    assert_state_close(g["state"], o["state"])
    vio = O.OracleVIO(frv["cam_cfg"], frv["ext"], frv["vio_cfg"])
    ov = vio.update(frv["img"], frv["vis_pos"], w["warp_patch"], w["search_levels"], frv["inv_ref_expo"], frv["state_prior"], frv["state_prior"])
    assert gv["total_iters"] == ov["total_iters"]
    assert_state_close(gv["state"], ov["state"], rot_tol=1e-8, pos_tol=1e-8, cov_tol=1e-6, rest_tol=1e-8)
    print(f"MULTI_GPU_OK mode={mode} tuning={tuning} world={world} lio_iters={g['iters']} M={g['M'].tolist()} vio_iters={gv['total_iters']}")
ctx.close()
dist.barrier(); dist.destroy_process_group()
