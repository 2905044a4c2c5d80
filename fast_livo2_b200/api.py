"""ctypes binding of the C ABI in include/esikf_b200.h (libesikf_b200.so).

This is plumbing for tests and bench.py — the product is the CUDA library. There is no CPU
fallback: importing works anywhere, but creating a context without the compiled library or
without an sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESIKF_LIB") or os.path.join(_HERE, "libesikf_b200.so")  # ESIKF_LIB: A/B builds of the same sources (measurement)
STATE_DOUBLES = 386


class EsikfError(RuntimeError):
    pass


class LioCfgC(C.Structure):
    _fields_ = [("voxel_size", C.c_double), ("sigma_num", C.c_double), ("dept_err", C.c_double), ("beam_err", C.c_double),
                ("max_layer", C.c_int32), ("max_iterations", C.c_int32)]


class ExtrinsicsC(C.Structure):
    _fields_ = [("extR", C.c_double * 9), ("extT", C.c_double * 3), ("Rcl", C.c_double * 9), ("Pcl", C.c_double * 3)]


class CameraC(C.Structure):
    _fields_ = [("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("pad_", C.c_int32), ("fx", C.c_double),
                ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("d", C.c_double * 5)]


class VioCfgC(C.Structure):
    _fields_ = [("img_point_cov", C.c_double), ("patch_pyrimid_level", C.c_int32), ("max_iterations", C.c_int32),
                ("exposure_estimate_en", C.c_int32), ("inverse_composition_en", C.c_int32)]


class LioStatsC(C.Structure):
    _fields_ = [("iters", C.c_int32), ("effct_feat_num", C.c_int32 * 8), ("converged", C.c_int32 * 8), ("pad_", C.c_int32),
                ("total_residual", C.c_double * 8), ("HTH", C.c_double * (8 * 36)), ("HTz", C.c_double * (8 * 6)),
                ("solution", C.c_double * (8 * 19))]


class VioStatsC(C.Structure):
    _fields_ = [("total_iters", C.c_int32), ("iters_per_level", C.c_int32 * 8), ("accepted_per_level", C.c_int32 * 8),
                ("pad_", C.c_int32), ("error_trace", C.c_float * 64), ("HTH", C.c_double * (64 * 49)), ("HTz", C.c_double * (64 * 7)),
                ("solution", C.c_double * (64 * 19))]


_lib = None


def load_library():
    """Load libesikf_b200.so; raises EsikfError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EsikfError(f"{LIB_PATH} is missing — run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, dp, fp, ip, i64p, u8p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(
        C.c_int64), C.POINTER(C.c_uint8)
    lib.esikf_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.esikf_destroy.argtypes = [vp]
    lib.esikf_last_error.argtypes = [vp]
    lib.esikf_last_error.restype = C.c_char_p
    lib.esikf_stream.argtypes = [vp]
    lib.esikf_stream.restype = vp
    lib.esikf_synchronize.argtypes = [vp]
    lib.esikf_launch_count.argtypes = [vp]
    lib.esikf_launch_count.restype = C.c_int64
    lib.esikf_set_solve_mode.argtypes = [vp, C.c_int]
    lib.esikf_set_loop_mode.argtypes = [vp, C.c_int]
    lib.esikf_set_tuning.argtypes = [vp, C.c_uint32]
    lib.esikf_set_extrinsics.argtypes = [vp, C.POINTER(ExtrinsicsC)]
    lib.esikf_map_upload.argtypes = [vp, i64p, ip, ip, C.c_int32, vp, C.c_int32, C.c_double]
    lib.esikf_map_patch.argtypes = [vp, ip, vp, C.c_int32]
    lib.esikf_map_device_init.argtypes = [vp, vp]
    lib.esikf_map_device_build.argtypes = [vp, vp]
    lib.esikf_map_device_update.argtypes = [vp, vp]
    lib.esikf_map_device_update_points.argtypes = [vp, vp, vp, C.c_int32]
    lib.esikf_map_device_stats.argtypes = [vp, vp]
    lib.esikf_map_device_slide.argtypes = [vp, vp, vp]
    lib.esikf_map_device_download.argtypes = [vp, vp, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
    lib.esikf_lio_fetch_normals.argtypes = [vp, vp]
    lib.esikf_lio_set_scan.argtypes = [vp, vp, C.c_int32]
    lib.esikf_lio_run.argtypes = [vp, vp, vp, C.POINTER(LioCfgC)]
    lib.esikf_lio_fetch.argtypes = [vp, vp, C.POINTER(LioStatsC), vp, vp, vp]
    lib.esikf_lio_update.argtypes = [vp, vp, C.c_int32, vp, vp, C.POINTER(LioCfgC), vp, C.POINTER(LioStatsC), vp, vp, vp]
    lib.esikf_lio_fetch_point_cov.argtypes = [vp, vp, vp]
    lib.esikf_vio_set_camera.argtypes = [vp, C.POINTER(CameraC), C.POINTER(VioCfgC)]
    lib.esikf_vio_set_image.argtypes = [vp, vp, C.c_int32, C.c_int32]
    lib.esikf_vio_set_patches.argtypes = [vp, vp, vp, vp, vp, C.c_int32]
    lib.esikf_vio_run.argtypes = [vp, vp, vp]
    lib.esikf_vio_fetch.argtypes = [vp, vp, C.POINTER(VioStatsC), vp]
    lib.esikf_vio_update.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, vp, C.POINTER(VioStatsC), vp]
    lib.esikf_vio_get_image_patch.argtypes = [vp, vp, C.c_int32, C.c_int32, vp]
    lib.esikf_vio_set_ref_images.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_int32, C.c_int32]
    lib.esikf_vio_warp_patches.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32]
    lib.esikf_vio_warp_affine.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.esikf_vio_set_inverse_refs.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.esikf_comm_unique_id.argtypes = [C.c_char_p]
    lib.esikf_comm_init.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p]
    lib.esikf_comm_rank.argtypes = [vp, ip, ip]
    lib.esikf_peer_export.argtypes = [vp, C.c_char_p]
    lib.esikf_peer_attach.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p]
    lib.esikf_shard_range.argtypes = [C.c_int32, C.c_int32, C.c_int32, ip, ip]
    lib.esikf_profile_kernel.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, fp]
    lib.esikf_set_kernel_timing.argtypes = [vp, C.c_int32]
    lib.esikf_set_phase_stamps.argtypes = [vp, C.c_int32]
    lib.esikf_get_phase_stamps.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.esikf_get_kernel_timing.argtypes = [vp, fp, fp, fp, fp]
    _lib = lib
    return lib


TUNE_STAGE_LDG = 1
TUNE_VIO_TMA = 2  # esikf_set_tuning flag: stage LIO plane records with __ldg copies instead of cp.async.bulk (measurement variant)
DEFAULT_LOOP_MODE = 2  # esikf_set_loop_mode: 2 (alias 1) persistent kernel per update, 0 per-iteration launches

class MapCfgC(C.Structure):
    _fields_ = [("voxel_size", C.c_double), ("min_eigen_value", C.c_double), ("dept_err", C.c_double), ("beam_err", C.c_double), ("max_layer", C.c_int32),
                ("max_points_num", C.c_int32), ("layer_init_num", C.c_int32 * 8), ("pad", C.c_int32), ("root_capacity", C.c_int64), ("node_capacity", C.c_int64),
                ("record_capacity", C.c_int64), ("point_capacity", C.c_int64)]


class MapStatsC(C.Structure):
    _fields_ = [("roots", C.c_int32), ("nodes", C.c_int32), ("records", C.c_int32), ("touched_roots", C.c_int32), ("errors", C.c_int32), ("pad", C.c_int32),
                ("pool_points", C.c_int64)]


EXPORTED_SYMBOLS = [
    "esikf_create", "esikf_destroy", "esikf_last_error", "esikf_stream", "esikf_synchronize", "esikf_host_alloc", "esikf_host_free", "esikf_launch_count", "esikf_set_solve_mode", "esikf_set_loop_mode", "esikf_set_tuning",
    "esikf_set_extrinsics", "esikf_set_lidar_extrinsics", "esikf_map_upload", "esikf_map_patch", "esikf_map_device_init", "esikf_map_device_build", "esikf_map_device_update", "esikf_map_device_update_points", "esikf_map_device_slide", "esikf_map_device_stats", "esikf_map_device_download", "esikf_lio_fetch_normals", "esikf_lio_set_scan", "esikf_lio_run", "esikf_lio_fetch",
    "esikf_lio_update", "esikf_lio_fetch_point_cov", "esikf_vio_set_camera", "esikf_vio_set_image", "esikf_vio_set_patches",
    "esikf_vio_run", "esikf_vio_fetch", "esikf_vio_update", "esikf_vio_get_image_patch", "esikf_vio_set_ref_images",
    "esikf_vio_warp_patches", "esikf_vio_warp_affine", "esikf_vio_set_inverse_refs", "esikf_comm_unique_id", "esikf_comm_init", "esikf_comm_rank", "esikf_shard_range", "esikf_peer_export", "esikf_peer_attach", "esikf_profile_kernel", "esikf_set_kernel_timing", "esikf_get_kernel_timing", "esikf_set_phase_stamps", "esikf_get_phase_stamps",
]


def _addr(a):
    """Raw address of a numpy array or torch tensor (host memory)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (pinned host memory in bench.py)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def lio_cfg_c(cfg) -> LioCfgC:
    return LioCfgC(cfg.voxel_size, cfg.sigma_num, cfg.dept_err, cfg.beam_err, cfg.max_layer, cfg.max_iterations)


class Context:
    """One ESIKF context = one CUDA device + stream (one process per GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.esikf_create(C.byref(h), device)
        if rc != 0:
            raise EsikfError(f"esikf_create(device={device}) failed with status {rc} (no sm_100 CUDA device?) — no CPU fallback")
        self.h = h
        self.n_pts = 0
        self.n_patches = 0
        self.levels = 4

    def close(self):
        if getattr(self, "h", None):
            self.lib.esikf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EsikfError(f"status {rc}: {self.lib.esikf_last_error(self.h).decode()}")

    @property
    def stream(self):
        return self.lib.esikf_stream(self.h)

    def synchronize(self):
        self._ck(self.lib.esikf_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.esikf_launch_count(self.h))

    def set_solve_mode(self, mode):
        self._ck(self.lib.esikf_set_solve_mode(self.h, mode))

    def set_loop_mode(self, mode):
        self._ck(self.lib.esikf_set_loop_mode(self.h, mode))

    def set_tuning(self, flags):
        """OR of the TUNE_* measurement variants; 0 = none."""
        self._ck(self.lib.esikf_set_tuning(self.h, flags))

    def set_extrinsics(self, ext):
        e = ExtrinsicsC()
        e.extR[:] = _c(ext.extR, np.float64).reshape(9)
        e.extT[:] = _c(ext.extT, np.float64)
        e.Rcl[:] = _c(ext.Rcl, np.float64).reshape(9)
        e.Pcl[:] = _c(ext.Pcl, np.float64)
        self._ck(self.lib.esikf_set_extrinsics(self.h, C.byref(e)))

    # ------------------------------------------------------------------ map
    def map_upload(self, vmap, voxel_size):
        k, f, c, p = _c(vmap["keys"], np.int64), _c(vmap["first"], np.int32), _c(vmap["count"], np.int32), np.ascontiguousarray(vmap["planes"])
        assert p.dtype.itemsize == 256
        self._ck(self.lib.esikf_map_upload(self.h, k.ctypes.data_as(C.POINTER(C.c_int64)), f.ctypes.data_as(C.POINTER(C.c_int32)),
                                           c.ctypes.data_as(C.POINTER(C.c_int32)), len(f), p.ctypes.data, len(p), float(voxel_size)))

    def map_patch(self, ids, planes):
        ids = _c(ids, np.int32)
        planes = np.ascontiguousarray(planes)
        self._ck(self.lib.esikf_map_patch(self.h, ids.ctypes.data_as(C.POINTER(C.c_int32)), planes.ctypes.data, len(ids)))

    # ------------------------------------------------------------------ device-resident map
    def map_device_init(self, cfg, root_capacity=0, node_capacity=0, record_capacity=0, point_capacity=0):
        """cfg: synthetic.LioCfg (voxel_size, min_eigen_value, dept_err, beam_err, max_layer, max_points_num, layer_init_num)."""
        m = MapCfgC()
        m.voxel_size, m.min_eigen_value, m.dept_err, m.beam_err = cfg.voxel_size, cfg.min_eigen_value, cfg.dept_err, cfg.beam_err
        m.max_layer, m.max_points_num = cfg.max_layer, cfg.max_points_num
        for k in range(8):
            m.layer_init_num[k] = cfg.layer_init_num[min(k, len(cfg.layer_init_num) - 1)]
        m.root_capacity, m.node_capacity, m.record_capacity, m.point_capacity = root_capacity, node_capacity, record_capacity, point_capacity
        self._ck(self.lib.esikf_map_device_init(self.h, C.byref(m)))

    def map_device_build(self, state):
        self._ck(self.lib.esikf_map_device_build(self.h, _c(state, np.float64).ctypes.data))

    def map_device_update(self, state=None):
        self._ck(self.lib.esikf_map_device_update(self.h, None if state is None else _c(state, np.float64).ctypes.data))

    def map_device_update_points(self, point_w, var):
        pw, v = _c(point_w, np.float64).reshape(-1, 3), _c(var, np.float64).reshape(-1, 9)
        self._ck(self.lib.esikf_map_device_update_points(self.h, pw.ctypes.data, v.ctypes.data, len(pw)))

    def map_device_slide(self, key_min=None, key_max=None):
        lo = None if key_min is None else _c(key_min, np.int64)
        hi = None if key_max is None else _c(key_max, np.int64)
        self._ck(self.lib.esikf_map_device_slide(self.h, None if lo is None else lo.ctypes.data, None if hi is None else hi.ctypes.data))

    def map_device_stats(self):
        st = MapStatsC()
        self._ck(self.lib.esikf_map_device_stats(self.h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in MapStatsC._fields_ if f != "pad"}

    def map_device_download(self):
        from .synthetic import PLANE_DTYPE

        nr, npl = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.esikf_map_device_download(self.h, None, None, None, 0, None, 0, C.byref(nr), C.byref(npl)))
        keys, first, count = np.zeros((max(nr.value, 1), 3), np.int64), np.zeros(max(nr.value, 1), np.int32), np.zeros(max(nr.value, 1), np.int32)
        planes = np.zeros(max(npl.value, 1), PLANE_DTYPE)
        self._ck(self.lib.esikf_map_device_download(self.h, keys.ctypes.data, first.ctypes.data, count.ctypes.data, len(first), planes.ctypes.data, len(planes),
                                                    C.byref(nr), C.byref(npl)))
        return dict(keys=keys[:nr.value], first=first[:nr.value], count=count[:nr.value], planes=planes[:npl.value])

    def lio_fetch_normals(self):
        out = np.zeros((self.n_pts, 3), np.float64)
        self._ck(self.lib.esikf_lio_fetch_normals(self.h, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ LIO
    def lio_set_scan(self, pts):
        """pts: (n,3) float32 numpy array or pinned torch tensor (kept alive by the caller until the copy is done)."""
        n = int(pts.shape[0])
        self._ck(self.lib.esikf_lio_set_scan(self.h, _addr(pts), n))
        self.n_pts = n

    def lio_run(self, state_in, state_prop, cfg):
        self._cfg_keep = lio_cfg_c(cfg) if not isinstance(cfg, LioCfgC) else cfg
        self._ck(self.lib.esikf_lio_run(self.h, _addr(state_in), _addr(state_prop), C.byref(self._cfg_keep)))

    def lio_fetch(self, per_point=True, state_out=None, match=None, normal=None, dis=None):
        """Outputs may be caller-provided (e.g. pinned torch tensors); otherwise numpy arrays are allocated."""
        out = np.zeros(STATE_DOUBLES) if state_out is None else state_out
        st = LioStatsC()
        n = self.n_pts
        if per_point:
            match = np.zeros(n, np.int32) if match is None else match
            normal = np.zeros(n, np.int32) if normal is None else normal
            dis = np.zeros(n, np.float32) if dis is None else dis
        self._ck(self.lib.esikf_lio_fetch(self.h, _addr(out), C.byref(st), _addr(match), _addr(normal), _addr(dis)))
        return self._lio_result(out, st, match, normal, dis)

    @staticmethod
    def _lio_result(out, st, match, normal, dis):
        it = st.iters
        return dict(state=out, match_plane=match, normal_plane=normal, dis_to_plane=dis, iters=it, M=np.array(st.effct_feat_num[:it]),
                    total_residual=np.array(st.total_residual[:it]), HTH=np.array(st.HTH[:]).reshape(8, 6, 6)[:it],
                    HTz=np.array(st.HTz[:]).reshape(8, 6)[:it], solution=np.array(st.solution[:]).reshape(8, 19)[:it],
                    converged=np.array(st.converged[:it]))

    def lio_update(self, pts, state_in, state_prop, cfg):
        """One-shot host-buffer call (esikf_lio_update): H2D + loop + D2H."""
        pts = _c(pts, np.float32)
        n = len(pts)
        self.n_pts = n
        out = np.zeros(STATE_DOUBLES)
        st = LioStatsC()
        match, normal, dis = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        cfgc = lio_cfg_c(cfg)
        si, sp = _c(state_in, np.float64), _c(state_prop, np.float64)
        self._ck(self.lib.esikf_lio_update(self.h, pts.ctypes.data, n, si.ctypes.data, sp.ctypes.data, C.byref(cfgc), out.ctypes.data, C.byref(st),
                                           match.ctypes.data, normal.ctypes.data, dis.ctypes.data))
        return self._lio_result(out, st, match, normal, dis)

    def lio_update_into(self, pts, state_in, state_prop, cfg, state_out, match, normal, dis):
        """esikf_lio_update with caller-owned (e.g. pinned) input AND output buffers: one C call per LIO tick."""
        n = int(pts.shape[0])
        self.n_pts = n
        st = LioStatsC()
        cfgc = lio_cfg_c(cfg) if not isinstance(cfg, LioCfgC) else cfg
        self._ck(self.lib.esikf_lio_update(self.h, _addr(pts), n, _addr(state_in), _addr(state_prop), C.byref(cfgc), _addr(state_out), C.byref(st),
                                           _addr(match), _addr(normal), _addr(dis)))
        return st.iters

    def vio_update_into(self, img, pos, warp_patch, search_levels, inv_expo, state_in, state_prop, state_out, errors):
        n = int(pos.shape[0])
        self.n_patches = n
        st = VioStatsC()
        self._ck(self.lib.esikf_vio_update(self.h, _addr(img), int(img.shape[1]), int(img.shape[0]), _addr(pos), _addr(warp_patch), _addr(search_levels),
                                           _addr(inv_expo), n, _addr(state_in), _addr(state_prop), _addr(state_out), C.byref(st), _addr(errors)))
        return st.total_iters

    def lio_fetch_point_cov(self):
        bc = np.zeros((self.n_pts, 3, 3))
        cm = np.zeros((self.n_pts, 3, 3))
        self._ck(self.lib.esikf_lio_fetch_point_cov(self.h, bc.ctypes.data, cm.ctypes.data))
        return bc, cm

    # ------------------------------------------------------------------ VIO
    def vio_set_camera(self, cam, vio):
        c = CameraC(cam.model, cam.width, cam.height, 0, cam.fx, cam.fy, cam.cx, cam.cy)
        c.d[:] = list(cam.d)
        v = VioCfgC(vio.img_point_cov, vio.levels, vio.max_iterations, int(vio.exposure_estimate_en), int(getattr(vio, "inverse_composition_en", False)))
        self._ck(self.lib.esikf_vio_set_camera(self.h, C.byref(c), C.byref(v)))
        self.levels = vio.levels

    def vio_set_image(self, img):
        h, w = int(img.shape[0]), int(img.shape[1])
        self._ck(self.lib.esikf_vio_set_image(self.h, _addr(img), w, h))

    def vio_set_patches(self, pos, warp_patch, search_levels, inv_expo):
        n = int(pos.shape[0])
        self._ck(self.lib.esikf_vio_set_patches(self.h, _addr(pos), _addr(warp_patch), _addr(search_levels), _addr(inv_expo), n))
        self.n_patches = n

    def vio_run(self, state_in, state_prop):
        self._ck(self.lib.esikf_vio_run(self.h, _addr(state_in), _addr(state_prop)))

    def vio_fetch(self, errors=True, state_out=None, err=None):
        out = np.zeros(STATE_DOUBLES) if state_out is None else state_out
        st = VioStatsC()
        if errors and err is None:
            err = np.zeros(self.n_patches, np.float32)
        self._ck(self.lib.esikf_vio_fetch(self.h, _addr(out), C.byref(st), _addr(err)))
        return self._vio_result(out, st, err)

    @staticmethod
    def _vio_result(out, st, err):
        return dict(state=out, errors=err, total_iters=st.total_iters, iters_per_level=np.array(st.iters_per_level[:]),
                    accepted_per_level=np.array(st.accepted_per_level[:]), error_trace=np.array(st.error_trace[:]).reshape(8, 8),
                    HTH=np.array(st.HTH[:]).reshape(8, 8, 7, 7), HTz=np.array(st.HTz[:]).reshape(8, 8, 7),
                    solution=np.array(st.solution[:]).reshape(8, 8, 19))

    def vio_update(self, img, pos, warp_patch, search_levels, inv_expo, state_in, state_prop):
        img = _c(img, np.uint8)
        pos, wp = _c(pos, np.float64), _c(warp_patch, np.float32)
        sl, ie = _c(search_levels, np.int32), _c(inv_expo, np.float64)
        n = len(pos)
        self.n_patches = n
        out = np.zeros(STATE_DOUBLES)
        st = VioStatsC()
        err = np.zeros(n, np.float32)
        si, sp = _c(state_in, np.float64), _c(state_prop, np.float64)
        self._ck(self.lib.esikf_vio_update(self.h, img.ctypes.data, img.shape[1], img.shape[0], pos.ctypes.data, wp.ctypes.data, sl.ctypes.data,
                                           ie.ctypes.data, n, si.ctypes.data, sp.ctypes.data, out.ctypes.data, C.byref(st), err.ctypes.data))
        return self._vio_result(out, st, err)

    def vio_get_image_patch(self, pc, level):
        pc = _c(pc, np.float64).reshape(-1, 2)
        out = np.zeros((len(pc), 64), np.float32)
        self._ck(self.lib.esikf_vio_get_image_patch(self.h, pc.ctypes.data, len(pc), level, out.ctypes.data))
        return out

    def vio_set_ref_images(self, imgs):
        imgs = [_c(im, np.uint8) for im in imgs]
        self._ref_keep = imgs
        arr = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
        self._ck(self.lib.esikf_vio_set_ref_images(self.h, arr, len(imgs), imgs[0].shape[1], imgs[0].shape[0]))

    def vio_warp_patches(self, ref_idx, px_ref, pos_w, normal_w, T_ref, T_cur, keep_on_device=False):
        n = len(px_ref)
        ref_idx = _c(ref_idx, np.int32)
        px_ref, pos_w, normal_w = _c(px_ref, np.float64), _c(pos_w, np.float64), _c(normal_w, np.float64)
        T_ref = _c(T_ref, np.float64).reshape(n, 12)
        T_cur = _c(T_cur, np.float64).reshape(12)
        A = np.zeros((n, 2, 2))
        sl = np.zeros(n, np.int32)
        wp = np.zeros((n, 64 * self.levels), np.float32)
        self._ck(self.lib.esikf_vio_warp_patches(self.h, n, ref_idx.ctypes.data, px_ref.ctypes.data, pos_w.ctypes.data, normal_w.ctypes.data,
                                                 T_ref.ctypes.data, T_cur.ctypes.data, A.ctypes.data, sl.ctypes.data, wp.ctypes.data,
                                                 int(keep_on_device)))
        if keep_on_device:
            self.n_patches = n
        return dict(A_cur_ref=A, search_levels=sl, warp_patch=wp)

    # ------------------------------------------------------------------ multi-GPU / measurement
    def vio_set_inverse_refs(self, ref_img_index, ref_px, ref_f, ref_R, ref_pos):
        """Reference-feature data of the inverse-compositional variant (esikf_vio_set_inverse_refs); images via vio_set_ref_images."""
        idx = _c(ref_img_index, np.int32)
        n = len(idx)
        px, f, R, pos = _c(ref_px, np.float64), _c(ref_f, np.float64), _c(np.asarray(ref_R).reshape(n, 9), np.float64), _c(ref_pos, np.float64)
        self._ck(self.lib.esikf_vio_set_inverse_refs(self.h, n, idx.ctypes.data, px.ctypes.data, f.ctypes.data, R.ctypes.data, pos.ctypes.data))

    def vio_warp_affine(self, ref_idx, px_ref, A_cur_ref, search_levels):
        """warpAffine alone (esikf_vio_warp_affine): caller-provided 2x2 matrices and search levels -> (n, levels*64) float32."""
        n = len(px_ref)
        ref_idx, sl = _c(ref_idx, np.int32), _c(search_levels, np.int32)
        px_ref, A = _c(px_ref, np.float64), _c(A_cur_ref, np.float64).reshape(n, 4)
        wp = np.zeros((n, 64 * self.levels), np.float32)
        self._ck(self.lib.esikf_vio_warp_affine(self.h, n, ref_idx.ctypes.data, px_ref.ctypes.data, A.ctypes.data, sl.ctypes.data, wp.ctypes.data))
        return wp

    def comm_init(self, rank, nranks, unique_id: bytes):
        self._ck(self.lib.esikf_comm_init(self.h, rank, nranks, unique_id))

    def set_phase_stamps(self, enable):
        self._ck(self.lib.esikf_set_phase_stamps(self.h, int(enable)))

    def get_phase_stamps(self):
        out = np.zeros(800, np.uint64)
        self._ck(self.lib.esikf_get_phase_stamps(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        self.debug_stamps = out[576:640].copy()
        self.cta_stamps = out[640:].copy()
        return out[:576].reshape(72, 8)

    def set_kernel_timing(self, enable):
        self._ck(self.lib.esikf_set_kernel_timing(self.h, int(enable)))

    def get_kernel_timing(self):
        a, b, c, d = (np.zeros(8, np.float32), np.zeros(8, np.float32), np.zeros(64, np.float32), np.zeros(64, np.float32))
        f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        self._ck(self.lib.esikf_get_kernel_timing(self.h, f(a), f(b), f(c), f(d)))
        return dict(lio_residual_ms=a, lio_solve_ms=b, vio_patch_ms=c, vio_solve_ms=d)

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.esikf_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, rank, nranks, handles):
        """handles: list of the 64-byte IPC handles of every rank, in rank order."""
        self._ck(self.lib.esikf_peer_attach(self.h, rank, nranks, b"".join(handles)))

    def profile_kernel(self, which, arg=0, reps=20, flush_l2=True):
        ms = C.c_float(0)
        self._ck(self.lib.esikf_profile_kernel(self.h, which, arg, reps, int(flush_l2), C.byref(ms)))
        return float(ms.value)


def shard_range(n, rank, nranks):
    lib = load_library()
    b, c = C.c_int32(0), C.c_int32(0)
    if lib.esikf_shard_range(n, rank, nranks, C.byref(b), C.byref(c)) != 0:
        raise EsikfError("esikf_shard_range: bad argument")
    return b.value, c.value


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.esikf_comm_unique_id(buf)
    if rc != 0:
        raise EsikfError(f"esikf_comm_unique_id failed: {rc}")
    return buf.raw


def pack_T(Rcw, Pcw):
    return np.concatenate([np.asarray(Rcw, np.float64).reshape(9), np.asarray(Pcw, np.float64).reshape(3)])
