// The call sequence of LIVMapper's LIO + VIO ticks (src/LIVMapper.cpp:351-428, 300-320) through the shim classes of
// fast_livo2_b200/csrc/fl2_shim.hpp — a compile-checked companion to INTEGRATION.md. The voxel map is DEVICE-RESIDENT here:
// BuildVoxelMap on the first frame, StateEstimation + UpdateVoxelMap() every tick; no plane, point list or key crosses PCIe
// and no host-side flatten / diff / patch runs between ticks (the host-owned form — SyncDeviceMap() after the reference's own
// UpdateVoxelMap — stays available). The scene is a floor patch and a trivial image: it shows WHICH calls
// happen WHEN and who owns what, not a meaningful estimate. Needs a B200 to run; without one the managers report an error.
//   g++ -std=c++17 -Iinclude examples/tick_loop.cpp -Lfast_livo2_b200 -lfl2_shim -lesikf_b200 -Wl,-rpath,$PWD/fast_livo2_b200 -o tick_loop
#include <cstdio>
#include <vector>

#include "../fast_livo2_b200/csrc/fl2_shim.hpp"

using namespace fl2b200;

static std::vector<PointXYZ> floor_scan(int n, float jitter) {
  std::vector<PointXYZ> pts(n);
  for (int i = 0; i < n; i++) {
    const float u = (float)(i % 40) * 0.05f, v = (float)(i / 40) * 0.05f;
    pts[i] = {u - 1.0f, v - 1.0f, -1.0f + jitter * (float)((i * 37) % 11 - 5) * 0.001f};
  }
  return pts;
}

int main() {
  // ---- node start-up (LIVMapper::initializeComponents): configuration; voxel_map stays empty — the device owns the map
  VoxelMapConfig cfg;
  cfg.max_voxel_size_ = 0.5, cfg.max_layer_ = 2, cfg.max_iterations_ = 5, cfg.sigma_num_ = 3.0, cfg.dept_err_ = 0.02, cfg.beam_err_ = 0.05;
  cfg.planner_threshold_ = 0.0025, cfg.max_points_num_ = 50, cfg.layer_init_num_ = {5, 5, 5, 5, 5};
  cfg.device_root_capacity_ = 1 << 16;
  VoxelMap voxel_map;
  VoxelMapManager voxelmap_manager(cfg, voxel_map, /*device*/ 0);
  if (voxelmap_manager.last_status_) {
    std::printf("no usable device (%s) — the managers never fall back to the CPU\n", voxelmap_manager.last_error_.c_str());
    return 0;
  }
  voxelmap_manager.EnableDeviceMap();
  voxelmap_manager.lazy_point_lists_ = true;             // pv_list_ & co. only when somebody reads them (MaterializePointLists)
  VIOManager vio_manager(voxelmap_manager.context());    // shares the device context and stream
  StatesGroup _state, state_propagat;                    // LIVMapper members; the managers hold pointers / copies like the reference
  vio_manager.state = &_state, vio_manager.state_propagat = &state_propagat;
  vio_manager.cam.model = 0, vio_manager.cam.width = 64, vio_manager.cam.height = 48;
  vio_manager.cam.fx = vio_manager.cam.fy = 60.0, vio_manager.cam.cx = 32.0, vio_manager.cam.cy = 24.0;
  vio_manager.patch_pyrimid_level = 2;
  vio_manager.initializeVIO();

  // ---- first LiDAR frame (LIVMapper.cpp:356-366): the map is built from the scan at the initial pose
  voxelmap_manager.feats_down_body_ = floor_scan(1600, 1.0f);
  voxelmap_manager.feats_down_size_ = (int)voxelmap_manager.feats_down_body_.size();
  voxelmap_manager.state_ = _state;                      // :257
  voxelmap_manager.BuildVoxelMap();                      // :364  (on the device)
  std::printf("BuildVoxelMap: status %d\n", voxelmap_manager.last_status_);

  for (int tick = 1; tick <= 3; tick++) {
    // ---- LIO tick (LIVMapper::handleLIO) after IMU propagation and down-sampling
    voxelmap_manager.feats_down_body_ = floor_scan(1600, 1.0f + 0.1f * tick);
    voxelmap_manager.feats_down_size_ = (int)voxelmap_manager.feats_down_body_.size();
    voxelmap_manager.state_ = _state;                    // :257
    voxelmap_manager.StateEstimation(state_propagat);    // :370
    _state = voxelmap_manager.state_;                    // :371
    voxelmap_manager.UpdateVoxelMap();                   // :413-424 in one call: world points, covariances, octree update, refits
    std::printf("tick %d LIO: status %d, effective points %d\n", tick, voxelmap_manager.last_status_, voxelmap_manager.effct_feat_num_);

    // ---- VIO tick (LIVMapper::handleVIO -> processFrame -> computeJacobianAndUpdateEKF, vio.cpp:1810)
    std::vector<uint8_t> pixels(64 * 48, 128);
    GrayImage img{pixels.data(), 64, 48};
    SubSparseMap submap;                                 // filled by retrieveFromVisualSparseMap on the host (vio.cpp:352-782)
    submap.voxel_points_pos.push_back(V3D());
    submap.voxel_points_pos[0][0] = 0.0, submap.voxel_points_pos[0][1] = 0.0, submap.voxel_points_pos[0][2] = 2.0;
    submap.warp_patch.push_back(std::vector<float>(64 * vio_manager.patch_pyrimid_level, 128.0f));
    submap.search_levels.push_back(0);
    submap.inv_expo_list.push_back(1.0);
    vio_manager.visual_submap = &submap;
    vio_manager.total_points = 1;
    state_propagat = _state;
    vio_manager.computeJacobianAndUpdateEKF(img);
    std::printf("tick %d VIO: status %d, patch error %.3f\n", tick, vio_manager.last_status_, submap.errors.empty() ? -1.0f : submap.errors[0]);
  }
  // somebody wants the per-point lists of the last tick after all (e.g. publishing the effective points)
  voxelmap_manager.MaterializePointLists();
  std::printf("pv_list_ %zu entries, ptpl_list_ %zu\n", voxelmap_manager.pv_list_.size(), voxelmap_manager.ptpl_list_.size());
  return 0;
}
