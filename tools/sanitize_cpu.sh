#!/bin/bash
# Address / undefined-behaviour sanitizer pass over the CPU code (oracle restatement + the shim's CPU-only entry points).
# Builds instrumented copies under /tmp/asan and drives them from Python with the sanitizer runtimes preloaded; prints the
# results of the calls and any sanitizer report (none expected). Leak detection is off: the interpreter and torch leak by design.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/asan
mkdir -p $OUT
CXX=/usr/bin/g++
FLAGS="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared"
(cd $ROOT/oracle && $CXX $FLAGS -ffp-contract=off -fopenmp -o $OUT/liborc_parity.so orc_lio.cpp orc_vio.cpp orc_capi.cpp)
$CXX $FLAGS -fopenmp -o $OUT/libfl2_shim.so $ROOT/fast_livo2_b200/csrc/fl2_shim.cpp -L$ROOT/fast_livo2_b200 -lesikf_b200 -Wl,-rpath,$ROOT/fast_livo2_b200
cd $ROOT
LD_PRELOAD="$($CXX -print-file-name=libasan.so) $($CXX -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 python - <<'PY'
import ctypes as C, sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_bind as O
O.ORACLE_DIR = "/tmp/asan"
from fast_livo2_b200 import synthetic as S
fr = S.make_frame(seed=11, n_pts=3000, n_map=100_000, n_patches=120, scene_scale=0.4)
lio = O.OracleLIO(fr["lio_cfg"], fr["ext"]); lio.set_map(fr["map"])
r = lio.state_estimation(fr["pts"], fr["state_prior"], fr["state_prior"])
w = O.oracle_warp_patches(fr, r["state"])
vio = O.OracleVIO(fr["cam_cfg"], fr["ext"], fr["vio_cfg"])
n = len(fr["vis_pos"])
v = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], fr["inv_ref_expo"], r["state"], r["state"])
vio.set_inverse_refs(**O.inverse_refs_from_frame(fr)); vio.set_inverse(True)
vi = vio.update(fr["img"], fr["vis_pos"], w["warp_patch"], w["search_levels"], np.ones(n), r["state"], r["state"])
print("oracle: LIO iters", r["iters"], "VIO iters", v["total_iters"], "inverse VIO iters", vi["total_iters"])
shim = C.CDLL("/tmp/asan/libfl2_shim.so")
m = fr["map"]
k, f, c, p = (np.ascontiguousarray(m["keys"], dtype=np.int64), np.ascontiguousarray(m["first"], dtype=np.int32), np.ascontiguousarray(m["count"], dtype=np.int32),
              np.ascontiguousarray(m["planes"]))
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ko, fo, co, po = np.zeros_like(k), np.zeros_like(f), np.zeros_like(c), np.zeros_like(p)
rc = shim.fl2_shim_flatten_roundtrip(vp(k), vp(f), vp(c), len(f), vp(p), len(p), C.c_double(float(fr["lio_cfg"].voxel_size)), int(fr["lio_cfg"].max_layer),
                                     vp(ko), vp(fo), vp(co), vp(po))
ids = np.zeros(len(p), np.int32); pb = p.copy(); pb["d"][3] += 1
print("shim: flatten round trip rc", rc, "diff", shim.fl2_shim_diff(vp(k), vp(f), vp(c), len(f), vp(p), len(p), vp(k), vp(f), vp(c), len(f), vp(pb), len(pb), vp(ids)))
PY
