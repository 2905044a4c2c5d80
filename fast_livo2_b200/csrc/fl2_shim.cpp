// fl2_shim.cpp — see fl2_shim.hpp. Host C++ only (no CUDA here): fills the reference-shaped member fields from / into
// the packed buffers of the C ABI (libesikf_b200.so, resolved at load time through the dynamic linker).
#include "fl2_shim.hpp"

#include <omp.h>

#include <chrono>

#include <cmath>
#include <cstring>

namespace fl2b200 {

// Host-side packing of the per-tick inputs into the page-locked staging buffers runs on a few threads (a num_threads clause, so
// nothing leaks into the process like the reference's omp_set_num_threads calls do): at 100 k points + 2 k patches the
// single-threaded copies were 0.29 ms of a 0.82 ms tick pair; 4 threads (the reference's MP_PROC_NUM) brought the pair to
// 0.71 ms, 8 are used since the copies are memory-bound and the GPU hosts have the cores.
#define FL2_PACK_THREADS 8
static void par_memcpy(void *dst, const void *src, size_t bytes) {
  const size_t chunk = 64 * 1024;
  const long n_chunks = (long)((bytes + chunk - 1) / chunk);
  if (n_chunks <= 2) {
    memcpy(dst, src, bytes);
    return;
  }
#pragma omp parallel for num_threads(FL2_PACK_THREADS) schedule(static)
  for (long c = 0; c < n_chunks; c++) {
    const size_t off = (size_t)c * chunk;
    memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, bytes - off < chunk ? bytes - off : chunk);
  }
}

StatesGroup::StatesGroup() {
  for (int i = 0; i < 361; i++) cov[i] = 0.0;
  for (int i = 0; i < 19; i++) cov[i * 19 + i] = 0.01;   // INIT_COV, include/common_lib.h:31,137
  cov[6 * 19 + 6] = 0.00001;                              // :138
  for (int i = 10; i < 19; i++) cov[i * 19 + i] = 0.00001;  // :139
}
void StatesGroup::pack(double *s) const {
  memcpy(s, rot_end.m, 9 * sizeof(double));
  memcpy(s + 9, pos_end.v, 3 * sizeof(double));
  s[12] = inv_expo_time;
  memcpy(s + 13, vel_end.v, 3 * sizeof(double));
  memcpy(s + 16, bias_g.v, 3 * sizeof(double));
  memcpy(s + 19, bias_a.v, 3 * sizeof(double));
  memcpy(s + 22, gravity.v, 3 * sizeof(double));
  memcpy(s + 25, cov, 361 * sizeof(double));
}
void StatesGroup::unpack(const double *s) {
  memcpy(rot_end.m, s, 9 * sizeof(double));
  memcpy(pos_end.v, s + 9, 3 * sizeof(double));
  inv_expo_time = s[12];
  memcpy(vel_end.v, s + 13, 3 * sizeof(double));
  memcpy(bias_g.v, s + 16, 3 * sizeof(double));
  memcpy(bias_a.v, s + 19, 3 * sizeof(double));
  memcpy(gravity.v, s + 22, 3 * sizeof(double));
  memcpy(cov, s + 25, 361 * sizeof(double));
}

pointWithVar::pointWithVar() {
  for (int i = 0; i < 9; i++) var_nostate.m[i] = body_var.m[i] = var.m[i] = point_crossmat.m[i] = 0.0;
}

VoxelOctoTree::~VoxelOctoTree() {
  for (int i = 0; i < 8; i++) delete leaves_[i];
  delete plane_ptr_;
}

static void flatten_node(const VoxelOctoTree *node, int layer, int max_layer, int path, FlatVoxelMap &out) {
  if (node->plane_ptr_ && node->plane_ptr_->is_plane_) {
    const VoxelPlane &p = *node->plane_ptr_;
    esikf_plane f;
    memset(&f, 0, sizeof(f));
    for (int k = 0; k < 3; k++) f.center[k] = p.center_[k], f.normal[k] = p.normal_[k];
    int t = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) f.plane_var[t++] = p.plane_var_[i * 6 + j];
    f.d = p.d_;
    f.radius = p.radius_;
    f.layer = layer;
    f.path = path;
    out.planes.push_back(f);
    out.plane_src.push_back(&p);
    return;
  }
  if (layer < max_layer)
    for (int l = 0; l < 8; l++)
      if (node->leaves_[l] != nullptr) flatten_node(node->leaves_[l], layer + 1, max_layer, path | (l << (3 * layer)), out);
}

bool FlattenVoxelMap(const VoxelMap &map, const VoxelMapConfig &cfg, FlatVoxelMap &out, std::string *err) {
  out.keys.clear(), out.first.clear(), out.count.clear(), out.planes.clear(), out.plane_src.clear();
  const float vs = (float)cfg.max_voxel_size_;
  for (const auto &kv : map) {
    const VoxelOctoTree *root = kv.second;
    if (!root) continue;
    // the device derives the root centre / quarter length from the key (src/voxel_map.cpp:578-581): check the host agrees
    const double c[3] = {(0.5 + kv.first.x) * vs, (0.5 + kv.first.y) * vs, (0.5 + kv.first.z) * vs};
    for (int k = 0; k < 3; k++)
      if (root->voxel_center_[k] != c[k] || root->quater_length_ != vs / 4) {
        if (err) *err = "root voxel centre / quarter length does not follow (0.5 + key) * voxel_size";
        return false;
      }
    out.keys.push_back(kv.first.x), out.keys.push_back(kv.first.y), out.keys.push_back(kv.first.z);
    out.first.push_back((int32_t)out.planes.size());
    flatten_node(root, 0, cfg.max_layer_, 0, out);
    out.count.push_back((int32_t)out.planes.size() - out.first.back());
  }
  return true;
}

bool DiffFlatVoxelMaps(const FlatVoxelMap &synced, const FlatVoxelMap &now, std::vector<int32_t> &changed_ids) {
  changed_ids.clear();
  if (synced.keys != now.keys || synced.first != now.first || synced.count != now.count || synced.planes.size() != now.planes.size()) return false;
  for (size_t i = 0; i < now.planes.size(); i++) {
    // layer / path are part of the record: a plane that moved inside its octree changes them and is patched like a refit
    if (memcmp(&synced.planes[i], &now.planes[i], sizeof(esikf_plane)) != 0) changed_ids.push_back((int32_t)i);
  }
  return true;
}

VoxelMapManager::VoxelMapManager(VoxelMapConfig &config_setting, VoxelMap &voxel_map, int device) : config_setting_(config_setting), voxel_map_(voxel_map) {
  last_status_ = esikf_create(&ctx_, device);
  if (last_status_ != 0) last_error_ = "esikf_create failed (no sm_100 CUDA device?) — there is no CPU fallback";
}
VoxelMapManager::~VoxelMapManager() { esikf_destroy(ctx_); }

void VoxelMapManager::SyncDeviceMap() {
  if (!ctx_) return;
  if (device_map_) {  // the device owns the map: nothing to mirror
    map_synced_ = true;
    return;
  }
  std::string err;
  FlatVoxelMap now;
  if (!FlattenVoxelMap(voxel_map_, config_setting_, now, &err)) {
    last_status_ = ESIKF_ERR_ARG, last_error_ = err;
    return;
  }
  std::vector<int32_t> ids;
  if (device_has_map_ && DiffFlatVoxelMaps(flat_, now, ids)) {
    // same roots and candidate lists as on the device: only the refitted records travel
    std::vector<esikf_plane> recs(ids.size());
    for (size_t k = 0; k < ids.size(); k++) recs[k] = now.planes[ids[k]];
    last_status_ = ids.empty() ? 0 : esikf_map_patch(ctx_, ids.data(), recs.data(), (int32_t)ids.size());
    last_sync_patched_ = (int)ids.size();
  } else {
    last_status_ = esikf_map_upload(ctx_, now.keys.data(), now.first.data(), now.count.data(), (int32_t)now.first.size(), now.planes.data(),
                                    (int32_t)now.planes.size(), config_setting_.max_voxel_size_);
    last_sync_patched_ = -1;
  }
  if (last_status_) {
    last_error_ = esikf_last_error(ctx_);
    device_has_map_ = false;  // unknown device state: the next sync uploads everything
  } else {
    flat_.keys.swap(now.keys), flat_.first.swap(now.first), flat_.count.swap(now.count), flat_.planes.swap(now.planes), flat_.plane_src.swap(now.plane_src);
    device_has_map_ = true;
  }
  map_synced_ = (last_status_ == 0);
}

void VoxelMapManager::EnableDeviceMap() {
  if (!ctx_) return;
  esikf_map_cfg m;
  memset(&m, 0, sizeof(m));
  m.voxel_size = config_setting_.max_voxel_size_, m.min_eigen_value = config_setting_.planner_threshold_;
  m.dept_err = config_setting_.dept_err_, m.beam_err = config_setting_.beam_err_;
  m.max_layer = config_setting_.max_layer_, m.max_points_num = config_setting_.max_points_num_;
  const std::vector<int> &lin = config_setting_.layer_init_num_;
  for (int k = 0; k < 8; k++) m.layer_init_num[k] = lin.empty() ? 5 : lin[k < (int)lin.size() ? k : (int)lin.size() - 1];
  m.root_capacity = config_setting_.device_root_capacity_;
  if ((last_status_ = esikf_map_device_init(ctx_, &m)) != 0) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  device_map_ = true, map_synced_ = true, device_has_map_ = true;
}

// include/voxel_map.h:231 / src/voxel_map.cpp:532-591 with feats_down_world_ = transformLidar(state_) (LIVMapper.cpp:360)
void VoxelMapManager::BuildVoxelMap() {
  if (!ctx_ || !device_map_) {
    last_status_ = ESIKF_ERR_STATE, last_error_ = "BuildVoxelMap: EnableDeviceMap first (a host-owned voxel_map_ is built by the reference's own code)";
    return;
  }
  const int n = feats_down_size_;
  float *pts = st_pts_.get((size_t)n * 3 + 4);
  double *sbuf = st_state_.get(3 * ESIKF_STATE_DOUBLES);
  if (!pts || !sbuf) {
    last_status_ = ESIKF_ERR_CUDA, last_error_ = "pinned staging allocation failed";
    return;
  }
  if (n) par_memcpy(pts, &feats_down_body_[0].x, (size_t)n * 12);
  state_.pack(sbuf);
  if ((last_status_ = esikf_set_lidar_extrinsics(ctx_, extR_.m, extT_.v)) == 0 && (last_status_ = esikf_lio_set_scan(ctx_, pts, n)) == 0)
    last_status_ = esikf_map_device_build(ctx_, sbuf);
  if (last_status_) last_error_ = esikf_last_error(ctx_);
}

// LIVMapper.cpp:413-424: pv_list_[i].point_w / .var with the posterior, then UpdateVoxelMap(pv_list_) — on the device
void VoxelMapManager::UpdateVoxelMap() {
  if (!ctx_ || !device_map_) {
    last_status_ = ESIKF_ERR_STATE, last_error_ = "UpdateVoxelMap(): EnableDeviceMap first";
    return;
  }
  double *sbuf = st_state_.get(3 * ESIKF_STATE_DOUBLES);
  state_.pack(sbuf);  // _state == voxelmap_manager->state_ at this point of the tick (LIVMapper.cpp:371); 3 kB
  if ((last_status_ = esikf_map_device_update(ctx_, sbuf)) != 0) last_error_ = esikf_last_error(ctx_);
}

// src/voxel_map.cpp:924-948: nothing until the sensor moved sliding_thresh, then roots further than half_map_size voxels go
void VoxelMapManager::mapSliding() {
  if (!ctx_ || !device_map_) {
    last_status_ = ESIKF_ERR_STATE, last_error_ = "mapSliding: EnableDeviceMap first (a host-owned voxel_map_ slides in the reference's own code)";
    return;
  }
  double d2 = 0;
  for (int k = 0; k < 3; k++) d2 += (position_last_[k] - last_slide_position[k]) * (position_last_[k] - last_slide_position[k]);
  if (sqrt(d2) < config_setting_.sliding_thresh) return;
  last_slide_position = position_last_;
  int64_t lo[3], hi[3];
  for (int j = 0; j < 3; j++) {
    float loc = (float)(position_last_[j] / config_setting_.max_voxel_size_);  // :938 (double quotient narrowed to float)
    if (loc < 0) loc -= 1.0;
    lo[j] = (int64_t)loc - config_setting_.half_map_size, hi[j] = (int64_t)loc + config_setting_.half_map_size;
  }
  if ((last_status_ = esikf_map_device_slide(ctx_, lo, hi)) != 0) last_error_ = esikf_last_error(ctx_);
}

// include/voxel_map.h:232 / src/voxel_map.cpp:609-641
void VoxelMapManager::UpdateVoxelMap(const std::vector<pointWithVar> &input_points) {
  if (!ctx_ || !device_map_) {
    last_status_ = ESIKF_ERR_STATE, last_error_ = "UpdateVoxelMap: EnableDeviceMap first";
    return;
  }
  const size_t n = input_points.size();
  double *buf = st_cov_.get(n * 12 + 2);
  if (!buf) {
    last_status_ = ESIKF_ERR_CUDA, last_error_ = "pinned staging allocation failed";
    return;
  }
  double *pw = buf, *var = buf + n * 3;
  for (size_t i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) pw[3 * i + k] = input_points[i].point_w[k];
    memcpy(var + 9 * i, input_points[i].var.m, 72);
  }
  if ((last_status_ = esikf_map_device_update_points(ctx_, pw, var, (int32_t)n)) != 0) last_error_ = esikf_last_error(ctx_);
}

// include/voxel_map.h:229 / src/voxel_map.cpp:338-511
void VoxelMapManager::StateEstimation(StatesGroup &state_propagat) {
  if (!ctx_) return;
  if (!map_synced_) SyncDeviceMap();
  if (!map_synced_) return;
  // lidar -> imu extrinsics only: the camera part belongs to the VIOManager sharing this context (it must survive the LIO tick)
  if ((last_status_ = esikf_set_lidar_extrinsics(ctx_, extR_.m, extT_.v)) != 0) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  const int n = feats_down_size_;
  esikf_lio_cfg cfg;
  cfg.voxel_size = config_setting_.max_voxel_size_, cfg.sigma_num = config_setting_.sigma_num_;
  cfg.dept_err = config_setting_.dept_err_, cfg.beam_err = config_setting_.beam_err_;
  cfg.max_layer = config_setting_.max_layer_, cfg.max_iterations = config_setting_.max_iterations_;
  // every buffer that crosses PCIe is staged in page-locked memory owned by the manager (grow-only, no allocation per tick)
  double *sbuf = st_state_.get(3 * ESIKF_STATE_DOUBLES);
  float *pts = st_pts_.get((size_t)n * 3 + 4), *dis = st_dis_.get((size_t)n + 1);
  int32_t *match = st_match_.get((size_t)n + 1), *normal = st_normal_.get((size_t)n + 1);
  if (!sbuf || !pts || !dis || !match || !normal) {
    last_status_ = ESIKF_ERR_CUDA, last_error_ = "pinned staging allocation failed";
    return;
  }
  double *sin = sbuf, *sprop = sbuf + ESIKF_STATE_DOUBLES, *sout = sbuf + 2 * ESIKF_STATE_DOUBLES;
  state_.pack(sin);
  state_propagat.pack(sprop);
  esikf_lio_stats stats;
  static_assert(sizeof(PointXYZ) == 12, "xyz float32");
  if (n) par_memcpy(pts, &feats_down_body_[0].x, (size_t)n * 12);
  last_status_ = esikf_lio_update(ctx_, n ? pts : nullptr, n, sin, sprop, &cfg, sout, &stats, match, normal, dis);
  if (last_status_) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  state_.unpack(sout);                 // _state = voxelmap_manager->state_  (LIVMapper.cpp:371)
  position_last_ = state_.pos_end;     // src/voxel_map.cpp:492
  effct_feat_num_ = stats.iters > 0 ? stats.effct_feat_num[stats.iters - 1] : 0;
  last_iters_ = stats.iters;
  lists_pending_ = fill_point_lists_;
  if (fill_point_lists_ && !lazy_point_lists_) MaterializePointLists();
}

// body_cov_list_ / cross_mat_list_ (LIVMapper.cpp:418-419), pv_list_ (:372) and ptpl_list_ (:446) of the last StateEstimation
void VoxelMapManager::MaterializePointLists() {
  if (!ctx_ || !lists_pending_) return;
  lists_pending_ = false;
  const int n = feats_down_size_;
  const int32_t *match = st_match_.data(), *normal = st_normal_.data();
  const float *dis = st_dis_.data();
  double *cov = st_cov_.get((size_t)n * 18 + 2);
  if (!cov) {
    last_status_ = ESIKF_ERR_CUDA, last_error_ = "pinned staging allocation failed";
    return;
  }
  const double *bc = cov, *cm = cov + (size_t)n * 9;
  if ((last_status_ = esikf_lio_fetch_point_cov(ctx_, cov, cov + (size_t)n * 9)) != 0) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  body_cov_list_.resize(n), cross_mat_list_.resize(n);
  pv_list_.resize(n);
  ptpl_list_.clear();
  const double *dev_normals = nullptr;
  if (device_map_) {  // no host VoxelPlane exists: pv.normal comes from the device records
    double *nb = st_normals_.get((size_t)n * 3 + 2);
    if (!nb || (last_status_ = esikf_lio_fetch_normals(ctx_, nb)) != 0) {
      last_error_ = nb ? esikf_last_error(ctx_) : "pinned staging allocation failed";
      if (!nb) last_status_ = ESIKF_ERR_CUDA;
      return;
    }
    dev_normals = nb;
  }
  const double *R = state_.rot_end.m;
  for (int i = 0; i < n; i++) {
    memcpy(body_cov_list_[i].m, &bc[(size_t)i * 9], 72);
    memcpy(cross_mat_list_[i].m, &cm[(size_t)i * 9], 72);
    pointWithVar &pv = pv_list_[i];
    pv = pointWithVar();  // the reference rebuilds pv_list_ from default-constructed elements every call (voxel_map.cpp:362-363)
    pv.point_b[0] = feats_down_body_[i].x, pv.point_b[1] = feats_down_body_[i].y, pv.point_b[2] = feats_down_body_[i].z;
    pv.body_var = body_cov_list_[i];
    // point_w / var are recomputed by the caller with the posterior state right after the call (LIVMapper.cpp:413-423);
    // point_w is provided here with the posterior pose for convenience.
    double pi[3], pw[3];
    for (int r = 0; r < 3; r++) pi[r] = extR_.m[3 * r] * pv.point_b[0] + extR_.m[3 * r + 1] * pv.point_b[1] + extR_.m[3 * r + 2] * pv.point_b[2] + extT_[r];
    for (int r = 0; r < 3; r++) pw[r] = R[3 * r] * pi[0] + R[3 * r + 1] * pi[1] + R[3 * r + 2] * pi[2] + state_.pos_end[r];
    for (int r = 0; r < 3; r++) pv.point_w[r] = (double)(float)pw[r];
    if (dev_normals) {
      for (int k = 0; k < 3; k++) pv.normal[k] = dev_normals[3 * (size_t)i + k];
      if (match[i] >= 0) {  // the plane's own fields stay on the device (esikf_map_device_download); what the tick produced is filled
        PointToPlane q;
        memset(q.plane_var_, 0, sizeof(q.plane_var_));
        q.point_b_ = pv.point_b, q.point_w_ = pv.point_w, q.normal_ = pv.normal, q.body_cov_ = pv.body_var;
        q.is_valid_ = true, q.dis_to_plane_ = dis[i];
        ptpl_list_.push_back(q);
      }
      continue;
    }
    if (normal[i] >= 0) pv.normal = flat_.plane_src[normal[i]]->normal_;  // pv.normal (voxel_map.cpp:744), zero if never matched
    if (match[i] >= 0) {
      const VoxelPlane &pl = *flat_.plane_src[match[i]];
      PointToPlane q;
      q.point_b_ = pv.point_b, q.point_w_ = pv.point_w, q.normal_ = pl.normal_, q.center_ = pl.center_;
      memcpy(q.plane_var_, pl.plane_var_, sizeof(q.plane_var_));
      q.body_cov_ = pv.body_var;
      q.layer_ = flat_.planes[match[i]].layer, q.d_ = pl.d_, q.is_valid_ = true, q.dis_to_plane_ = dis[i];
      ptpl_list_.push_back(q);
    }
  }
}

VIOManager::VIOManager(esikf_ctx *shared_ctx) : ctx_(shared_ctx) {}

void VIOManager::initializeVIO() {
  if (!ctx_) return;
  esikf_extrinsics ext;
  memcpy(ext.extR, extR.m, sizeof(ext.extR));
  memcpy(ext.extT, extT.v, sizeof(ext.extT));
  memcpy(ext.Rcl, Rcl.m, sizeof(ext.Rcl));
  memcpy(ext.Pcl, Pcl.v, sizeof(ext.Pcl));
  last_status_ = esikf_set_extrinsics(ctx_, &ext);
  esikf_vio_cfg cfg;
  cfg.img_point_cov = img_point_cov, cfg.patch_pyrimid_level = patch_pyrimid_level, cfg.max_iterations = max_iterations;
  cfg.exposure_estimate_en = exposure_estimate_en ? 1 : 0, cfg.inverse_composition_en = 0;
  if (!last_status_) last_status_ = esikf_vio_set_camera(ctx_, &cam, &cfg);
  if (last_status_) last_error_ = esikf_last_error(ctx_);
}

// include/vio.h:153 / src/vio.cpp:784-802
void VIOManager::computeJacobianAndUpdateEKF(const GrayImage &img) {
  if (!ctx_ || !state || !state_propagat || !visual_submap) return;
  if (total_points == 0) return;  // src/vio.cpp:786
  const int n = total_points, L = patch_pyrimid_level;
  double *pos = st_pos_.get((size_t)n * 3), *ie = st_ie_.get(n), *sbuf = st_state_.get(3 * ESIKF_STATE_DOUBLES);
  float *wp = st_wp_.get((size_t)n * 64 * L), *err = st_err_.get(n);
  int32_t *sl = st_sl_.get(n);
  uint8_t *im = st_img_.get((size_t)img.cols * img.rows);
  if (!pos || !ie || !sbuf || !wp || !err || !sl || !im) {
    last_status_ = ESIKF_ERR_CUDA, last_error_ = "pinned staging allocation failed";
    return;
  }
  const SubSparseMap &sub = *visual_submap;
#pragma omp parallel for num_threads(FL2_PACK_THREADS) schedule(static) if (n > 256)
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) pos[(size_t)3 * i + k] = sub.voxel_points_pos[i][k];
    memcpy(&wp[(size_t)i * 64 * L], sub.warp_patch[i].data(), sizeof(float) * 64 * L);
    sl[i] = visual_submap->search_levels[i];
    ie[i] = visual_submap->inv_expo_list[i];
  }
  par_memcpy(im, img.data, (size_t)img.cols * img.rows);
  double *sin = sbuf, *sprop = sbuf + ESIKF_STATE_DOUBLES, *sout = sbuf + 2 * ESIKF_STATE_DOUBLES;
  state->pack(sin);
  state_propagat->pack(sprop);
  esikf_vio_stats stats;
  patch_img_ = nullptr;  // the update installs its own current image
  last_status_ = esikf_vio_update(ctx_, im, img.cols, img.rows, pos, wp, sl, ie, n, sin, sprop, sout, &stats, err);
  if (last_status_) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  visual_submap->errors.assign(err, err + n);
  state->unpack(sout);
  last_total_iters_ = stats.total_iters;
}

// include/vio.h:151 / src/vio.cpp:203-225: writes patch_tmp[patch_size_total * level + row * patch_size + col]
void VIOManager::getImagePatch(const GrayImage &img, const double pc[2], float *patch_tmp, int level) {
  if (!ctx_ || !img.data || !pc || !patch_tmp) return;
  last_status_ = 0;
  if (img.data != patch_img_ || img.cols != patch_img_w_ || img.rows != patch_img_h_) {
    last_status_ = esikf_vio_set_image(ctx_, img.data, img.cols, img.rows);
    patch_img_ = last_status_ ? nullptr : img.data, patch_img_w_ = img.cols, patch_img_h_ = img.rows;
  }
  float out[64];
  if (!last_status_) last_status_ = esikf_vio_get_image_patch(ctx_, pc, 1, level, out);
  if (last_status_) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  memcpy(patch_tmp + 64 * level, out, sizeof(out));
}

// include/vio.h:161-162 / src/vio.cpp:292-318: writes patch[patch_size_total * pyramid_level + y * patch_size + x].
// level_ref is unused by the reference as well; halfpatch_size must be 4 (patch_size 8).
void VIOManager::warpAffine(const double A_cur_ref[4], const GrayImage &img_ref, const double px_ref[2], int level_ref, int search_level, int pyramid_level,
                            int halfpatch_size, float *patch) {
  (void)level_ref;
  if (!ctx_ || !img_ref.data || !A_cur_ref || !px_ref || !patch) return;
  if (halfpatch_size != 4 || pyramid_level < 0 || pyramid_level >= patch_pyrimid_level) {
    last_status_ = ESIKF_ERR_ARG;
    last_error_ = "warpAffine: halfpatch_size must be 4 and pyramid_level inside the pyramid";
    return;
  }
  last_status_ = 0;
  if (img_ref.data != patch_ref_img_ || img_ref.cols != patch_ref_w_ || img_ref.rows != patch_ref_h_) {
    const uint8_t *imgs[1] = {img_ref.data};
    last_status_ = esikf_vio_set_ref_images(ctx_, imgs, 1, img_ref.cols, img_ref.rows);
    patch_ref_img_ = last_status_ ? nullptr : img_ref.data, patch_ref_w_ = img_ref.cols, patch_ref_h_ = img_ref.rows;
  }
  std::vector<float> all((size_t)64 * patch_pyrimid_level);
  const int32_t idx = 0, sl = search_level;
  if (!last_status_) last_status_ = esikf_vio_warp_affine(ctx_, 1, &idx, px_ref, A_cur_ref, &sl, all.data());
  if (last_status_) {
    last_error_ = esikf_last_error(ctx_);
    return;
  }
  // a singular A leaves the patch untouched in the reference (vio.cpp:297-301); the kernel writes zeros into its own
  // scratch in that case, so only copy when A is invertible
  const double det = A_cur_ref[0] * A_cur_ref[3] - A_cur_ref[1] * A_cur_ref[2];
  const double inv_det = 1.0 / det;
  const float inv00 = (float)(A_cur_ref[3] * inv_det);  // same expression as the kernel's A_ref_cur(0, 0)
  if (inv00 != inv00) return;
  memcpy(patch + 64 * pyramid_level, all.data() + 64 * pyramid_level, sizeof(float) * 64);
}

}  // namespace fl2b200

// ---------------------------------------------------------------------------------------------------------------------
// C entry points for the parity tests (Python / ctypes): rebuild a pointer octree from flat arrays, run the shim on it.
using namespace fl2b200;

static void build_tree(VoxelMap &map, const VoxelMapConfig &cfg, const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots,
                       const esikf_plane *planes) {
  const float vs = (float)cfg.max_voxel_size_;
  for (int r = 0; r < n_roots; r++) {
    VOXEL_LOCATION loc(keys[3 * r], keys[3 * r + 1], keys[3 * r + 2]);
    VoxelOctoTree *root = new VoxelOctoTree;
    root->plane_ptr_ = new VoxelPlane;
    root->quater_length_ = vs / 4;
    root->voxel_center_[0] = (0.5 + loc.x) * vs, root->voxel_center_[1] = (0.5 + loc.y) * vs, root->voxel_center_[2] = (0.5 + loc.z) * vs;
    root->init_octo_ = true;
    map[loc] = root;
    for (int c = 0; c < count[r]; c++) {
      const esikf_plane &f = planes[first[r] + c];
      VoxelOctoTree *node = root;
      for (int l = 0; l < f.layer; l++) {
        const int leaf = (f.path >> (3 * l)) & 7;
        if (!node->leaves_[leaf]) {
          VoxelOctoTree *ch = new VoxelOctoTree;
          ch->plane_ptr_ = new VoxelPlane;
          ch->layer_ = l + 1;
          const int xyz[3] = {(leaf >> 2) & 1, (leaf >> 1) & 1, leaf & 1};
          for (int k = 0; k < 3; k++) ch->voxel_center_[k] = node->voxel_center_[k] + (2 * xyz[k] - 1) * node->quater_length_;
          ch->quater_length_ = node->quater_length_ / 2;
          ch->init_octo_ = true;
          node->leaves_[leaf] = ch;
        }
        node = node->leaves_[leaf];
      }
      VoxelPlane &p = *node->plane_ptr_;
      for (int k = 0; k < 3; k++) p.center_[k] = f.center[k], p.normal_[k] = f.normal[k];
      int t = 0;
      for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) p.plane_var_[i * 6 + j] = p.plane_var_[j * 6 + i] = f.plane_var[t++];
      p.d_ = f.d, p.radius_ = f.radius, p.is_plane_ = true, p.is_init_ = true;
    }
  }
}

extern "C" {

// CPU-only: flat arrays -> pointer octree -> FlattenVoxelMap -> flat arrays (sizes returned; buffers sized like the input).
int fl2_shim_flatten_roundtrip(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const esikf_plane *planes, int n_planes,
                               double voxel_size, int max_layer, int64_t *keys_out, int32_t *first_out, int32_t *count_out, esikf_plane *planes_out) {
  VoxelMapConfig cfg;
  cfg.max_voxel_size_ = voxel_size, cfg.max_layer_ = max_layer;
  VoxelMap map;
  build_tree(map, cfg, keys, first, count, n_roots, planes);
  FlatVoxelMap out;
  std::string err;
  const bool ok = FlattenVoxelMap(map, cfg, out, &err);
  for (auto &kv : map) delete kv.second;
  if (!ok || (int)out.first.size() != n_roots || (int)out.planes.size() != n_planes) return -1;
  memcpy(keys_out, out.keys.data(), out.keys.size() * sizeof(int64_t));
  memcpy(first_out, out.first.data(), out.first.size() * sizeof(int32_t));
  memcpy(count_out, out.count.data(), out.count.size() * sizeof(int32_t));
  memcpy(planes_out, out.planes.data(), out.planes.size() * sizeof(esikf_plane));
  return 0;
}

// GPU: VoxelMapManager::StateEstimation + VIOManager::computeJacobianAndUpdateEKF through the shim classes.
// vio arrays may be null (LIO only). Returns 0 or the first failing esikf_status.
int fl2_shim_run(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const esikf_plane *planes, int n_planes,
                 const esikf_lio_cfg *lcfg, const esikf_extrinsics *ext, const float *pts, int n, const double *state_in, const double *state_prop,
                 double *lio_state_out, int32_t *n_effective, int32_t *n_ptpl, double *normals_out /* n x 3 */, const esikf_camera *cam,
                 const esikf_vio_cfg *vcfg, const uint8_t *img, int n_patches, const double *pos, const float *warp_patch, const int32_t *search_levels,
                 const double *inv_expo, double *vio_state_out, float *errors_out) {
  (void)n_planes;
  VoxelMapConfig cfg;
  cfg.max_voxel_size_ = lcfg->voxel_size, cfg.max_layer_ = lcfg->max_layer, cfg.max_iterations_ = lcfg->max_iterations;
  cfg.beam_err_ = lcfg->beam_err, cfg.dept_err_ = lcfg->dept_err, cfg.sigma_num_ = lcfg->sigma_num;
  VoxelMap map;
  build_tree(map, cfg, keys, first, count, n_roots, planes);
  int rc = 0;
  {
    VoxelMapManager mgr(cfg, map, 0);
    if (mgr.last_status_) rc = mgr.last_status_;
    memcpy(mgr.extR_.m, ext->extR, 72);
    memcpy(mgr.extT_.v, ext->extT, 24);
    mgr.feats_down_body_.resize(n);
    memcpy(mgr.feats_down_body_.data(), pts, (size_t)n * 12);
    mgr.feats_down_size_ = n;
    mgr.state_.unpack(state_in);              // voxelmap_manager->state_ = _state  (LIVMapper.cpp:257)
    StatesGroup prop;
    prop.unpack(state_prop);
    if (!rc) {
      mgr.SyncDeviceMap();
      mgr.StateEstimation(prop);              // LIVMapper.cpp:370
      rc = mgr.last_status_;
    }
    if (!rc) {
      mgr.state_.pack(lio_state_out);
      *n_effective = mgr.effct_feat_num_;
      *n_ptpl = (int32_t)mgr.ptpl_list_.size();
      for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) normals_out[3 * i + k] = mgr.pv_list_[i].normal[k];
    }
    if (!rc && cam && n_patches > 0) {
      VIOManager vio(mgr.context());
      StatesGroup st = mgr.state_, stp = mgr.state_;
      SubSparseMap sub;
      const int L = vcfg->patch_pyrimid_level;
      for (int i = 0; i < n_patches; i++) {
        V3D p;
        for (int k = 0; k < 3; k++) p[k] = pos[3 * i + k];
        sub.voxel_points_pos.push_back(p);
        sub.warp_patch.emplace_back(warp_patch + (size_t)i * 64 * L, warp_patch + (size_t)(i + 1) * 64 * L);
        sub.search_levels.push_back(search_levels[i]);
        sub.inv_expo_list.push_back(inv_expo[i]);
      }
      vio.state = &st, vio.state_propagat = &stp, vio.visual_submap = &sub, vio.total_points = n_patches;
      vio.patch_pyrimid_level = L, vio.max_iterations = vcfg->max_iterations, vio.img_point_cov = vcfg->img_point_cov;
      vio.exposure_estimate_en = vcfg->exposure_estimate_en != 0;
      vio.cam = *cam;
      memcpy(vio.Rcl.m, ext->Rcl, 72), memcpy(vio.Pcl.v, ext->Pcl, 24), memcpy(vio.extR.m, ext->extR, 72), memcpy(vio.extT.v, ext->extT, 24);
      vio.initializeVIO();
      GrayImage im;
      im.data = img, im.cols = cam->width, im.rows = cam->height;
      if (!vio.last_status_) vio.computeJacobianAndUpdateEKF(im);
      rc = vio.last_status_;
      if (!rc) {
        st.pack(vio_state_out);
        for (int i = 0; i < n_patches; i++) errors_out[i] = sub.errors[i];
      }
    }
  }
  for (auto &kv : map) delete kv.second;
  return rc;
}

// CPU-only: DiffFlatVoxelMaps on two flat maps given as arrays (same root arrays for both unless keys_b is non-null).
// Returns -1 if the structure differs, else the number of changed plane ids written to ids_out (capacity n_planes_a).
int fl2_shim_diff(const int64_t *keys_a, const int32_t *first_a, const int32_t *count_a, int n_roots_a, const esikf_plane *planes_a, int n_planes_a,
                  const int64_t *keys_b, const int32_t *first_b, const int32_t *count_b, int n_roots_b, const esikf_plane *planes_b, int n_planes_b,
                  int32_t *ids_out) {
  FlatVoxelMap a, b;
  a.keys.assign(keys_a, keys_a + 3 * (size_t)n_roots_a), a.first.assign(first_a, first_a + n_roots_a), a.count.assign(count_a, count_a + n_roots_a);
  a.planes.assign(planes_a, planes_a + n_planes_a);
  b.keys.assign(keys_b, keys_b + 3 * (size_t)n_roots_b), b.first.assign(first_b, first_b + n_roots_b), b.count.assign(count_b, count_b + n_roots_b);
  b.planes.assign(planes_b, planes_b + n_planes_b);
  std::vector<int32_t> ids;
  if (!DiffFlatVoxelMaps(a, b, ids)) return -1;
  for (size_t k = 0; k < ids.size(); k++) ids_out[k] = ids[k];
  return (int)ids.size();
}

// GPU: map refresh through the shim. Builds the octree of map A, syncs (full upload), overwrites the plane fits with map B's
// (same structure: an UpdateVoxelMap that only refitted planes), syncs again (incremental: esikf_map_patch of the changed
// records) and runs StateEstimation. The caller compares with a run on a fresh manager holding map B.
int fl2_shim_resync_run(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const esikf_plane *planes_a, const esikf_plane *planes_b,
                        int n_planes, const esikf_lio_cfg *lcfg, const esikf_extrinsics *ext, const float *pts, int n, const double *state_in,
                        double *state_out, int32_t *n_patched) {
  (void)n_planes;
  VoxelMapConfig cfg;
  cfg.max_voxel_size_ = lcfg->voxel_size, cfg.max_layer_ = lcfg->max_layer, cfg.max_iterations_ = lcfg->max_iterations;
  cfg.beam_err_ = lcfg->beam_err, cfg.dept_err_ = lcfg->dept_err, cfg.sigma_num_ = lcfg->sigma_num;
  VoxelMap map;
  build_tree(map, cfg, keys, first, count, n_roots, planes_a);
  int rc = 0;
  {
    VoxelMapManager mgr(cfg, map, 0);
    rc = mgr.last_status_;
    memcpy(mgr.extR_.m, ext->extR, 72);
    memcpy(mgr.extT_.v, ext->extT, 24);
    if (!rc) {
      mgr.SyncDeviceMap();
      rc = mgr.last_status_;
    }
    if (!rc) {
      // the host map changes under the manager: same octree, new plane fits
      for (int r = 0; r < n_roots; r++) {
        VOXEL_LOCATION loc(keys[3 * r], keys[3 * r + 1], keys[3 * r + 2]);
        for (int c = 0; c < count[r]; c++) {
          const esikf_plane &f = planes_b[first[r] + c];
          VoxelOctoTree *node = map[loc];
          for (int l = 0; l < f.layer; l++) node = node->leaves_[(f.path >> (3 * l)) & 7];
          VoxelPlane &p = *node->plane_ptr_;
          for (int k = 0; k < 3; k++) p.center_[k] = f.center[k], p.normal_[k] = f.normal[k];
          int t = 0;
          for (int i = 0; i < 6; i++)
            for (int j = i; j < 6; j++) p.plane_var_[i * 6 + j] = p.plane_var_[j * 6 + i] = f.plane_var[t++];
          p.d_ = f.d, p.radius_ = f.radius;
        }
      }
      mgr.MarkMapDirty();
      mgr.SyncDeviceMap();
      rc = mgr.last_status_;
      *n_patched = mgr.last_sync_patched_;
    }
    if (!rc) {
      mgr.feats_down_body_.resize(n);
      memcpy(mgr.feats_down_body_.data(), pts, (size_t)n * 12);
      mgr.feats_down_size_ = n;
      mgr.fill_point_lists_ = false;
      mgr.state_.unpack(state_in);
      StatesGroup prop;
      prop.unpack(state_in);
      mgr.StateEstimation(prop);
      rc = mgr.last_status_;
      if (!rc) mgr.state_.pack(state_out);
    }
  }
  for (auto &kv : map) delete kv.second;
  return rc;
}

// GPU: VIOManager::getImagePatch / VIOManager::warpAffine through the shim class (one patch each).
int fl2_shim_patch_helpers(const esikf_camera *cam, const esikf_vio_cfg *vcfg, const uint8_t *img, int cols, int rows, const double *pc, int level,
                           float *patch_tmp /* levels*64, only `level` written */, const double *A_cur_ref, const uint8_t *img_ref, const double *px_ref,
                           int search_level, int pyramid_level, float *warp_patch /* levels*64, only `pyramid_level` written */) {
  esikf_ctx *ctx = nullptr;
  int rc = esikf_create(&ctx, 0);
  if (rc) return rc;
  {
    VIOManager vio(ctx);
    vio.cam = *cam;
    vio.patch_pyrimid_level = vcfg->patch_pyrimid_level, vio.max_iterations = vcfg->max_iterations, vio.img_point_cov = vcfg->img_point_cov;
    vio.exposure_estimate_en = vcfg->exposure_estimate_en != 0;
    vio.initializeVIO();
    rc = vio.last_status_;
    GrayImage g{img, cols, rows}, gr{img_ref, cols, rows};
    if (!rc) {
      vio.getImagePatch(g, pc, patch_tmp, level);
      rc = vio.last_status_;
    }
    if (!rc) {
      vio.warpAffine(A_cur_ref, gr, px_ref, 0, search_level, pyramid_level, 4, warp_patch);
      rc = vio.last_status_;
    }
  }
  esikf_destroy(ctx);
  return rc;
}

// ---- persistent session for the end-to-end measurement through the drop-in classes (bench.py e2e_shim): the managers,
// the host octree and the visual sub-map live across ticks like they do inside LIVMapper.
struct ShimSession {
  VoxelMapConfig cfg;
  VoxelMap map;
  VoxelMapManager *mgr = nullptr;
  VIOManager *vio = nullptr;
  SubSparseMap sub;
  StatesGroup st, stp;
  bool has_vio = false;
  esikf_camera cam{};
  esikf_vio_cfg vcfg{};
  long long manager_ns = 0;
  ~ShimSession() {
    delete vio;
    delete mgr;
    for (auto &kv : map) delete kv.second;
  }
};

void *fl2_shim_session_create(const int64_t *keys, const int32_t *first, const int32_t *count, int n_roots, const esikf_plane *planes, int n_planes,
                              const esikf_lio_cfg *lcfg, const esikf_extrinsics *ext, const esikf_camera *cam, const esikf_vio_cfg *vcfg, int device) {
  (void)n_planes;
  ShimSession *s = new ShimSession;
  s->cfg.max_voxel_size_ = lcfg->voxel_size, s->cfg.max_layer_ = lcfg->max_layer, s->cfg.max_iterations_ = lcfg->max_iterations;
  s->cfg.beam_err_ = lcfg->beam_err, s->cfg.dept_err_ = lcfg->dept_err, s->cfg.sigma_num_ = lcfg->sigma_num;
  build_tree(s->map, s->cfg, keys, first, count, n_roots, planes);
  s->mgr = new VoxelMapManager(s->cfg, s->map, device);
  if (s->mgr->last_status_) {
    delete s;
    return nullptr;
  }
  memcpy(s->mgr->extR_.m, ext->extR, 72);
  memcpy(s->mgr->extT_.v, ext->extT, 24);
  s->mgr->SyncDeviceMap();
  if (s->mgr->last_status_) {
    delete s;
    return nullptr;
  }
  if (cam && vcfg) {
    s->has_vio = true, s->cam = *cam, s->vcfg = *vcfg;
    s->vio = new VIOManager(s->mgr->context());
    VIOManager &vio = *s->vio;
    vio.state = &s->st, vio.state_propagat = &s->stp, vio.visual_submap = &s->sub;
    vio.patch_pyrimid_level = vcfg->patch_pyrimid_level, vio.max_iterations = vcfg->max_iterations, vio.img_point_cov = vcfg->img_point_cov;
    vio.exposure_estimate_en = vcfg->exposure_estimate_en != 0;
    vio.cam = *cam;
    memcpy(vio.Rcl.m, ext->Rcl, 72), memcpy(vio.Pcl.v, ext->Pcl, 24), memcpy(vio.extR.m, ext->extR, 72), memcpy(vio.extT.v, ext->extT, 24);
    vio.initializeVIO();
    if (vio.last_status_) {
      delete s;
      return nullptr;
    }
  }
  return s;
}

// One LIVMapper tick pair: LIO (LIVMapper.cpp:356-372) then VIO on the LIO posterior (:305, vio.cpp:1810). iters_out[0] = LIO
// iterations, [1] = VIO iterations.
int fl2_shim_session_step(void *h, const float *pts, int n, const double *state_in, const double *state_prop, const uint8_t *img, int n_patches, const double *pos,
                          const float *warp_patch, const int32_t *search_levels, const double *inv_expo, double *lio_state_out, double *vio_state_out, int32_t *iters_out) {
  ShimSession *s = static_cast<ShimSession *>(h);
  if (!s) return ESIKF_ERR_ARG;
  VoxelMapManager &mgr = *s->mgr;
  mgr.feats_down_body_.resize(n);
  if (n) memcpy(mgr.feats_down_body_.data(), pts, (size_t)n * 12);
  mgr.feats_down_size_ = n;
  mgr.state_.unpack(state_in);  // voxelmap_manager->state_ = _state  (LIVMapper.cpp:257)
  StatesGroup prop;
  prop.unpack(state_prop);
  const auto t_lio0 = std::chrono::steady_clock::now();
  mgr.StateEstimation(prop);    // LIVMapper.cpp:370
  s->manager_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_lio0).count();
  if (mgr.last_status_) return mgr.last_status_;
  mgr.state_.pack(lio_state_out);
  iters_out[0] = mgr.last_iters_, iters_out[1] = 0;
  if (s->has_vio && n_patches > 0) {
    VIOManager &vio = *s->vio;
    s->st = mgr.state_, s->stp = mgr.state_;
    SubSparseMap &sub = s->sub;
    const int L = s->vcfg.patch_pyrimid_level;
    if ((int)sub.voxel_points_pos.size() != n_patches) {  // the visual sub-map of the tick (retrieveFromVisualSparseMap's output)
      sub.voxel_points_pos.resize(n_patches), sub.warp_patch.resize(n_patches), sub.search_levels.resize(n_patches), sub.inv_expo_list.resize(n_patches);
    }
    for (int i = 0; i < n_patches; i++) {
      for (int k = 0; k < 3; k++) sub.voxel_points_pos[i][k] = pos[3 * i + k];
      sub.warp_patch[i].assign(warp_patch + (size_t)i * 64 * L, warp_patch + (size_t)(i + 1) * 64 * L);
      sub.search_levels[i] = search_levels[i];
      sub.inv_expo_list[i] = inv_expo[i];
    }
    vio.total_points = n_patches;
    GrayImage im;
    im.data = img, im.cols = s->cam.width, im.rows = s->cam.height;
    const auto t_vio0 = std::chrono::steady_clock::now();
    vio.computeJacobianAndUpdateEKF(im);
    s->manager_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_vio0).count();
    if (vio.last_status_) return vio.last_status_;
    s->st.pack(vio_state_out);
    iters_out[1] = vio.last_total_iters_;
  }
  return 0;
}

// mode 0: pv_list_ / ptpl_list_ / body_cov_list_ / cross_mat_list_ filled inside every StateEstimation (the reference's member
// contract); 1: filled lazily (MaterializePointLists — never, in the timed loop of a caller that does not read them); 2: off.
void fl2_shim_session_point_lists(void *h, int mode) {
  ShimSession *s = static_cast<ShimSession *>(h);
  if (!s) return;
  s->mgr->fill_point_lists_ = (mode != 2);
  s->mgr->lazy_point_lists_ = (mode == 1);
}
// number of entries the lists hold after materialising them (checks the lazy path in the tests)
int fl2_shim_session_materialize(void *h, int32_t *n_pv, int32_t *n_ptpl) {
  ShimSession *s = static_cast<ShimSession *>(h);
  if (!s) return ESIKF_ERR_ARG;
  s->mgr->MaterializePointLists();
  if (n_pv) *n_pv = (int32_t)s->mgr->pv_list_.size();
  if (n_ptpl) *n_ptpl = (int32_t)s->mgr->ptpl_list_.size();
  return s->mgr->last_status_;
}

// Device-resident map for the session: the map the session was created with is dropped, an empty device map takes its place and
// absorbs `pts` at `state` through BuildVoxelMap (first LiDAR frame, LIVMapper.cpp:356-366).
int fl2_shim_session_device_map(void *h, const float *pts, int n, const double *state, double min_eigen_value, int max_points_num, long long root_capacity) {
  ShimSession *s = static_cast<ShimSession *>(h);
  if (!s) return ESIKF_ERR_ARG;
  VoxelMapManager &mgr = *s->mgr;
  mgr.config_setting_.planner_threshold_ = min_eigen_value, mgr.config_setting_.max_points_num_ = max_points_num;
  mgr.config_setting_.device_root_capacity_ = root_capacity;
  mgr.EnableDeviceMap();
  if (mgr.last_status_) return mgr.last_status_;
  mgr.feats_down_body_.resize(n);
  if (n) memcpy(mgr.feats_down_body_.data(), pts, (size_t)n * 12);
  mgr.feats_down_size_ = n;
  mgr.state_.unpack(state);
  mgr.BuildVoxelMap();
  return mgr.last_status_;
}
// LIVMapper.cpp:413-424 after the LIO half of a tick: the device map absorbs the scan with the LIO posterior.
int fl2_shim_session_update_map(void *h) {
  ShimSession *s = static_cast<ShimSession *>(h);
  if (!s) return ESIKF_ERR_ARG;
  s->mgr->UpdateVoxelMap();
  return s->mgr->last_status_;
}

// nanoseconds the last step spent inside StateEstimation + computeJacobianAndUpdateEKF (the two calls LIVMapper makes; filling
// the managers' members from the harness' flat buffers — which LIVMapper already holds in that shape — is outside)
long long fl2_shim_session_manager_ns(void *h) {
  ShimSession *s = static_cast<ShimSession *>(h);
  return s ? s->manager_ns : 0;
}

void fl2_shim_session_destroy(void *h) { delete static_cast<ShimSession *>(h); }

}  // extern "C"
