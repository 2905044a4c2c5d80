"""Generate tests/golden/esikf_golden.npz: a small seeded LIO+VIO frame and the CPU oracle's outputs on it.

The reference ships no golden vectors for this path and cannot be built here, so these vectors pin the
ORACLE (regression) and let the GPU tests check the CUDA path without running the oracle. Regenerate with:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_bind as O  # noqa: E402
from fast_livo2_b200 import synthetic as S  # noqa: E402


def main():
    fr = S.make_frame(seed=21, n_pts=1500, n_map=60_000, n_patches=60, scene_scale=0.3)
    lio = O.OracleLIO(fr["lio_cfg"], fr["ext"])
    lio.set_map(fr["map"])
    o = lio.state_estimation(fr["pts"], fr["state_prior"], fr["state_prior"])
    w = O.oracle_warp_patches(fr, o["state"])
    vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
    v = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], o["state"], o["state"])
    m = fr["map"]
    np.savez_compressed(
        os.path.join(HERE, "esikf_golden.npz"),
        map_keys=m["keys"], map_first=m["first"], map_count=m["count"], map_planes=m["planes"].view(np.uint8).reshape(len(m["planes"]), 256),
        pts=fr["pts"], state_prior=fr["state_prior"], lio_cfg=fr["lio_cfg"].as_array(), vio_cfg=fr["vio_cfg"].as_array(), cam_cfg=fr["cam_cfg"].as_array(),
        extR=fr["ext"].extR, extT=fr["ext"].extT, Rcl=fr["ext"].Rcl, Pcl=fr["ext"].Pcl,
        lio_state=o["state"], lio_iters=o["iters"], lio_M=o["M"], lio_match=o["match_plane"], lio_normal=o["normal_plane"], lio_dis=o["dis_to_plane"],
        lio_HTH=o["HTH"], lio_HTz=o["HTz"],
        img=fr["img"], img_ref=fr["img_ref"], vis_pos=fr["vis_pos"], vis_normal=fr["vis_normal"], px_ref=fr["px_ref"], T_ref_R=fr["T_ref"][0], T_ref_t=fr["T_ref"][1],
        inv_ref_expo=fr["inv_ref_expo"], warp_patch=w["warp_patch"], search_levels=w["search_levels"], A_cur_ref=w["A_cur_ref"],
        vio_state=v["state"], vio_total_iters=v["total_iters"], vio_iters_per_level=v["iters_per_level"], vio_errors=v["errors"], vio_error_trace=v["error_trace"])
    print("written", os.path.getsize(os.path.join(HERE, "esikf_golden.npz")) // 1024, "KiB; LIO iters", o["iters"], "M", o["M"], "VIO iters", v["total_iters"])
    # inverse-compositional variant (vio/inverse_composition_en) on the same inputs: only the extra inputs and the outputs
    refs = O.inverse_refs_from_frame(fr)
    vio.set_inverse_refs(**refs)
    vio.set_inverse(True)
    n = len(fr["vis_pos"])
    vi = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], np.ones(n), o["state"], o["state"])
    Hinv = vio.precompute_reference_patches(fr["vis_pos"], 1)
    np.savez_compressed(os.path.join(HERE, "esikf_golden_inverse.npz"), ref_img_index=refs["ref_img_index"], ref_px=refs["ref_px"], ref_f=refs["ref_f"],
                        ref_R=refs["ref_R"], ref_pos=refs["ref_pos"], vio_state=vi["state"], vio_total_iters=vi["total_iters"],
                        vio_iters_per_level=vi["iters_per_level"], vio_accepted_per_level=vi["accepted_per_level"], vio_errors=vi["errors"],
                        vio_error_trace=vi["error_trace"], H_sub_inv_level1=Hinv[:8])
    print("inverse variant: VIO iters", vi["total_iters"], vi["iters_per_level"][:4])


if __name__ == "__main__":
    main()
