// Device-side common definitions of the B200 ESIKF update (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/esikf_b200.h"

namespace esikf {

// ---- packed state offsets (ESIKF_STATE_DOUBLES doubles, see include/esikf_b200.h)
enum { S_R = 0, S_P = 9, S_EXPO = 12, S_V = 13, S_BG = 16, S_BA = 19, S_G = 22, S_COV = 25, S_N = ESIKF_STATE_DOUBLES };

// ---- device hash of root voxels: open addressing, linear probing, 16-byte slots
struct HashSlot {
  unsigned long long key;  // packed VOXEL_LOCATION (21 bits / axis, biased), ~0 = empty
  uint32_t first;          // first candidate plane
  uint32_t count;          // number of candidate planes (DFS order)
};
#define ESIKF_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define ESIKF_KEY_BIAS (1 << 20)
#define ESIKF_KEY_RANGE (1 << 21)

__host__ __device__ inline bool key_in_range(long long x, long long y, long long z) {
  return x >= -ESIKF_KEY_BIAS && x < ESIKF_KEY_BIAS && y >= -ESIKF_KEY_BIAS && y < ESIKF_KEY_BIAS && z >= -ESIKF_KEY_BIAS &&
         z < ESIKF_KEY_BIAS;
}
__host__ __device__ inline unsigned long long pack_key(long long x, long long y, long long z) {
  return ((unsigned long long)(x + ESIKF_KEY_BIAS) << 42) | ((unsigned long long)(y + ESIKF_KEY_BIAS) << 21) |
         (unsigned long long)(z + ESIKF_KEY_BIAS);
}
__host__ __device__ inline uint32_t hash_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

// ---- information buffer produced by the residual kernels and consumed by the solve kernel.
// Layout (doubles): [0:64) D = sum_i a_i (w_i a_i)^T, 8x8 row-major, a = [H(6|7), z, 1|res]
//                   [64] count (matched points / n_meas)   [65] sum|d| (LIO)   [66..71] spare
enum { INFO_D = 0, INFO_COUNT = 64, INFO_ABS = 65, INFO_N = 72 };

// ---- loop-control block shared by the kernels of one update (device memory)
struct Ctrl {
  int stop;         // LIO: EKF_stop_flg ; VIO: whole update finished
  int iter;         // LIO: iterations executed so far ; VIO: total iterations
  int rematch_num;  // LIO (voxel_map.cpp:365,482)
  int level_done;   // VIO: EKF_end of the current level
  int level;        // VIO: level being processed
  int level_iter;   // VIO: iteration counter inside the level
  float last_error; // VIO (vio.cpp:1528)
  int has_G;        // VIO: G valid (at least one accepted update)
  unsigned int block_counter;  // last-block-done counter of the residual kernels
  int pad[7];
};


#ifdef __CUDACC__
// fp64 tensor-core contraction step: D(8x8) += A(8x4) * B(4x8), mma.sync.m8n8k4.f64 (SASS DMMA).
// Lane l holds A[l/4][l%4], B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Shared scratch of the fixed-order reduction used by both residual kernels.
template <int WARPS> struct ReduceSmem {
  double warpD[WARPS][66];
  double seg[3][INFO_N];
  int is_last;
};

// Combine every warp's 8x8 block (D0, D1 fragments) and scalar count into info[] :
// warp -> block (fixed warp order) -> grid (last block to finish sums the per-block partials in block order).
// abs_in_77: LIO keeps sum|d| in D[7][7]; it is moved to info[INFO_ABS] and D[7][7] zeroed.
template <int WARPS>
__device__ __forceinline__ void reduce_info(ReduceSmem<WARPS> &rs, double D0, double D1, double cnt, bool abs_in_77, double *partials,
                                            double *info, Ctrl *ctrl) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int g = lane >> 2, t = lane & 3;
    rs.warpD[warp][g * 8 + 2 * t] = D0;
    rs.warpD[warp][g * 8 + 2 * t + 1] = D1;
    if (lane == 0) rs.warpD[warp][64] = cnt;
  }
  __syncthreads();
  if (tid < 65) {
    double s = rs.warpD[0][tid];
#pragma unroll
    for (int w = 1; w < WARPS; w++) s += rs.warpD[w][tid];
    double *out = partials + (size_t)blockIdx.x * INFO_N;
    if (tid < 64) {
      if (abs_in_77 && tid == 63) {
        out[63] = 0.0;
        out[INFO_ABS] = s;
      } else {
        out[tid] = s;
        if (tid == 63) out[INFO_ABS] = 0.0;
      }
    } else {
      out[INFO_COUNT] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int prev = atomicAdd(&ctrl->block_counter, 1u);
    rs.is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!rs.is_last) return;
  __threadfence();
  const int nb = gridDim.x;
  if (tid < 3 * INFO_N) {
    const int e = tid % INFO_N, seg = tid / INFO_N;
    const int b0 = (nb * seg) / 3, b1 = (nb * (seg + 1)) / 3;
    double s = 0.0;
    if (e < 66)
      for (int b = b0; b < b1; b++) s += __ldcg(partials + (size_t)b * INFO_N + e);
    rs.seg[seg][e] = s;
  }
  __syncthreads();
  if (tid < INFO_N) info[tid] = (rs.seg[0][tid] + rs.seg[1][tid]) + rs.seg[2][tid];
  if (tid == 0) ctrl->block_counter = 0;
}
#endif

}  // namespace esikf
