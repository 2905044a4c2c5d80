// The call sequence of one LIO + VIO tick pair as LIVMapper would issue it (src/LIVMapper.cpp:351-428, 300-320) through the
// shim classes of fast_livo2_b200/csrc/fl2_shim.hpp — a compile-checked companion to INTEGRATION.md. The map here is one
// root voxel holding a single floor plane and the scan / image are trivial; it shows WHICH calls happen WHEN and who owns
// what, not a meaningful estimate. Needs a B200 to run; without one the managers report ESIKF_ERR_NO_DEVICE.
//   g++ -std=c++17 -Iinclude examples/tick_loop.cpp -Lfast_livo2_b200 -lfl2_shim -lesikf_b200 -Wl,-rpath,$PWD/fast_livo2_b200 -o tick_loop
#include <cstdio>
#include <vector>

#include "../fast_livo2_b200/csrc/fl2_shim.hpp"

using namespace fl2b200;

int main() {
  // ---- node start-up: configuration and the (host-owned) voxel map, as LIVMapper::initializeComponents does
  VoxelMapConfig cfg;
  cfg.max_voxel_size_ = 0.5, cfg.max_layer_ = 2, cfg.max_iterations_ = 5, cfg.sigma_num_ = 3.0, cfg.dept_err_ = 0.02, cfg.beam_err_ = 0.05;
  VoxelMap voxel_map;
  VOXEL_LOCATION loc(0, 0, -1);  // the voxel [0, 0.5) x [0, 0.5) x [-0.5, 0)
  VoxelOctoTree *root = new VoxelOctoTree;
  root->plane_ptr_ = new VoxelPlane;
  root->quater_length_ = (float)cfg.max_voxel_size_ / 4;
  root->voxel_center_[0] = 0.25, root->voxel_center_[1] = 0.25, root->voxel_center_[2] = -0.25;
  root->init_octo_ = true;
  VoxelPlane &pl = *root->plane_ptr_;
  pl.center_[0] = 0.25, pl.center_[1] = 0.25, pl.center_[2] = -0.1;
  pl.normal_[0] = 0, pl.normal_[1] = 0, pl.normal_[2] = 1;
  pl.d_ = 0.1f, pl.radius_ = 0.3f, pl.is_plane_ = true, pl.is_init_ = true;
  for (int i = 0; i < 6; i++) pl.plane_var_[i * 6 + i] = 1e-4;
  voxel_map[loc] = root;

  VoxelMapManager voxelmap_manager(cfg, voxel_map, /*device*/ 0);
  if (voxelmap_manager.last_status_) {
    std::printf("no usable device (%s) — the managers never fall back to the CPU\n", voxelmap_manager.last_error_.c_str());
    delete root;
    return 0;
  }
  VIOManager vio_manager(voxelmap_manager.context());  // shares the device context and stream
  StatesGroup _state, state_propagat;                  // LIVMapper members; the managers hold pointers / copies like the reference
  vio_manager.state = &_state, vio_manager.state_propagat = &state_propagat;
  vio_manager.cam.model = 0, vio_manager.cam.width = 64, vio_manager.cam.height = 48;
  vio_manager.cam.fx = vio_manager.cam.fy = 60.0, vio_manager.cam.cx = 32.0, vio_manager.cam.cy = 24.0;
  vio_manager.patch_pyrimid_level = 2;
  vio_manager.initializeVIO();

  // ---- LIO tick (LIVMapper::handleLIO): after IMU propagation and down-sampling
  voxelmap_manager.feats_down_body_ = {{0.10f, 0.10f, -0.12f}, {0.30f, 0.20f, -0.09f}, {0.20f, 0.40f, -0.11f}};
  voxelmap_manager.feats_down_size_ = (int)voxelmap_manager.feats_down_body_.size();
  voxelmap_manager.state_ = _state;                       // LIVMapper.cpp:257
  voxelmap_manager.SyncDeviceMap();                       // first call: full upload; later calls patch refitted planes only
  voxelmap_manager.StateEstimation(state_propagat);       // LIVMapper.cpp:370
  _state = voxelmap_manager.state_;                       // :371
  std::printf("LIO: status %d, effective points %d, pv_list_ %zu entries\n", voxelmap_manager.last_status_, voxelmap_manager.effct_feat_num_,
              voxelmap_manager.pv_list_.size());
  // ... UpdateVoxelMap(pv_list_) runs on the host here (LIVMapper.cpp:424) and changes voxel_map ...
  voxelmap_manager.MarkMapDirty();                        // the next StateEstimation refreshes the device mirror

  // ---- VIO tick (LIVMapper::handleVIO -> processFrame -> computeJacobianAndUpdateEKF, vio.cpp:1810)
  std::vector<uint8_t> pixels(64 * 48, 128);
  GrayImage img{pixels.data(), 64, 48};
  SubSparseMap submap;                                    // filled by retrieveFromVisualSparseMap on the host (vio.cpp:352-782)
  submap.voxel_points_pos.push_back(V3D());
  submap.voxel_points_pos[0][0] = 0.0, submap.voxel_points_pos[0][1] = 0.0, submap.voxel_points_pos[0][2] = 2.0;
  submap.warp_patch.push_back(std::vector<float>(64 * vio_manager.patch_pyrimid_level, 128.0f));
  submap.search_levels.push_back(0);
  submap.inv_expo_list.push_back(1.0);
  vio_manager.visual_submap = &submap;
  vio_manager.total_points = 1;
  state_propagat = _state;
  vio_manager.computeJacobianAndUpdateEKF(img);
  std::printf("VIO: status %d, patch error %.3f\n", vio_manager.last_status_, submap.errors.empty() ? -1.0f : submap.errors[0]);
  delete root;
  return 0;
}
