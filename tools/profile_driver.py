"""Small driver for ncu captures: builds the BASELINE config-2 frame and runs a few LIO+VIO updates."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_livo2_b200 import api, synthetic as S

n_pts = int(os.environ.get("N_PTS", 100000)); n_patch = int(os.environ.get("N_PATCH", 2000)); steps = int(os.environ.get("STEPS", 3))
fr = S.cached_frame(seed=0, n_pts=n_pts, n_map=int(os.environ.get("N_MAP", 1000000)), n_patches=n_patch)
ctx = api.Context(0)
if os.environ.get("ESIKF_LOOP") is not None:
    ctx.set_loop_mode(int(os.environ["ESIKF_LOOP"]))  # 0: per-iteration launches (profiles the stand-alone residual kernels)
ctx.set_extrinsics(fr["ext"]); ctx.map_upload(fr["map"], fr["lio_cfg"].voxel_size); ctx.vio_set_camera(fr["cam_cfg"], fr["vio_cfg"])
r = ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
st = S.unpack_state(r["state"])
ctx.vio_set_image(fr["img"]); ctx.vio_set_ref_images([fr["img_ref"]])
n = len(fr["vis_pos"])
w = ctx.vio_warp_patches(np.zeros(n, np.int32), fr["px_ref"], fr["vis_pos"], fr["vis_normal"], np.tile(api.pack_T(*fr["T_ref"]), (n, 1)),
                         api.pack_T(*S.camera_pose(fr["ext"], st["R"], st["p"])))
for _ in range(steps):
    r = ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
    v = ctx.vio_update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], r["state"], r["state"])
print("iters", r["iters"], v["total_iters"], "M", r["M"])
if os.environ.get("ESIKF_MAP"):  # device-resident map: build from the scan at the true pose, then LIO update + map update per step
    ctx.map_device_init(fr["lio_cfg"], root_capacity=1 << 19)
    ctx.lio_set_scan(fr["pts"])
    ctx.map_device_build(fr["state_true"])
    for _ in range(steps):
        r = ctx.lio_update(fr["pts"], fr["state_prior"], fr["state_prior"], fr["lio_cfg"])
        ctx.map_device_update()
    print("device map", ctx.map_device_stats(), "M", r["M"])
