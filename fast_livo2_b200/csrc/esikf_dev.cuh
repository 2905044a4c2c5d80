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
  int accepted_in_level;       // VIO: accepted updates of the current level (diagnostics)
  int pad[6];
};


#ifdef __CUDACC__
// Measurement-only fine-grained timestamps (ns) by CTA 0 / thread 0 into a 64-entry debug array (nullable).
__device__ __forceinline__ void dbg_stamp(unsigned long long *dbg, int k) {
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[k] = t;
  }
}

// fp64 tensor-core contraction step: D(8x8) += A(8x4) * B(4x8), mma.sync.m8n8k4.f64 (SASS DMMA).
// Lane l holds A[l/4][l%4], B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Fixed-order sum over the per-CTA partial blocks (entry-major [entry][block]) by the calling CTA (>= 14 warps, nb <= 160
// blocks): warp w owns entries w, w+nwarps, ...; lane l adds blocks l, l+32, ... in order, then a fixed xor-shuffle tree.
// Every load is issued before the first add, so the whole sum costs ONE L2 round trip.
#define SUM_MAXE 5
#define SUM_MAXC 5
__device__ __forceinline__ void sum_partials(const double *partials, int partial_stride, int nb, double *info, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[SUM_MAXE][SUM_MAXC];
#pragma unroll
  for (int k = 0; k < SUM_MAXE; k++) {
    const int e = warp + k * nwarps;
    const double *__restrict__ p = partials + (size_t)e * partial_stride;
#pragma unroll
    for (int c = 0; c < SUM_MAXC; c++) {
      const int b = lane + 32 * c;
      v[k][c] = (e < 66 && b < nb) ? __ldcg(p + b) : 0.0;
    }
  }
#pragma unroll
  for (int k = 0; k < SUM_MAXE; k++) {
    const int e = warp + k * nwarps;
    double s = v[k][0];
#pragma unroll
    for (int c = 1; c < SUM_MAXC; c++) s += v[k][c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0 && e < 66) info[e] = s;
  }
}

// Shared scratch of the fixed-order reduction used by both residual kernels.
template <int WARPS> struct ReduceSmem {
  double warpD[WARPS][66];
  int is_last;
};

// Combine every warp's 8x8 block (D0, D1 fragments) and scalar count into info[] :
// warp -> block (fixed warp order) -> grid. Per-block partials are stored entry-major ([entry][block]) so the last block
// to finish can sum each entry with coalesced loads: lane l adds blocks l, l+32, ... in order, then a fixed xor-shuffle
// tree. The order never depends on which block is last => bit-reproducible.
// abs_in_77: LIO keeps sum|d| in D[7][7]; it is moved to info[INFO_ABS] and D[7][7] zeroed.
template <int WARPS>
__device__ __forceinline__ void reduce_info(ReduceSmem<WARPS> &rs, double D0, double D1, double cnt, bool abs_in_77, double *partials,
                                            int partial_stride, double *info, Ctrl *ctrl) {
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
    int e = tid;
    if (tid == 64) e = INFO_COUNT;
    if (abs_in_77 && tid == 63) e = INFO_ABS;
    partials[(size_t)e * partial_stride + blockIdx.x] = s;
    if (tid == 63) partials[(size_t)(abs_in_77 ? 63 : INFO_ABS) * partial_stride + blockIdx.x] = 0.0;
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
  sum_partials(partials, partial_stride, gridDim.x, info, WARPS);
  if (tid >= 66 && tid < INFO_N) info[tid] = 0.0;
  if (tid == 0) ctrl->block_counter = 0;
}
#endif

}  // namespace esikf
