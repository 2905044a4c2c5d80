// Solve routines of the B200 ESIKF update: the Kalman gain, the boxplus state update, loop control and the final
// covariance update. One block: all threads stage P / info / poses in a single global round trip, one warp runs the
// m x m gain solve, all threads write back. The whole iteration loop runs on the device with no host round trip.
//
//   lio_solve_kernel : src/voxel_map.cpp:462-499  (K_1, G, solution, state_ += solution, convergence / rematch / (I-G)P)
//   vio_solve_kernel : src/vio.cpp:1636-1685 + :800 (error-gated accept / rollback, K_1, G, solution, final cov -= G cov)
//
// Gain: the reference computes K_1 = (H^T H + P^-1)^-1 with two 19x19 partial-pivot inversions and then only uses the
// first m (6 or 7) columns of K_1. Because H^T H is zero outside its leading m x m block A, the push-through identity gives
//     K_1[:, :m] = P[:, :m] (I_m + A P_mm)^-1
// exactly — an m x m solve with 19 right-hand sides, one per lane. solve_mode 1 keeps the literal double inversion (Gauss-
// Jordan with partial pivoting in shared memory) for parity checks.
#include <float.h>
#include "esikf_dev.cuh"

namespace esikf {

struct SolveArgs {
  double *state;        // current iterate (device, packed) — updated in place
  const double *prop;   // state_propagat
  const double *info;   // reduced information buffer
  Ctrl *ctrl;
  int max_iterations;
  int solve_mode;
  // LIO
  esikf_lio_stats *lio_stats;
  // VIO
  esikf_vio_stats *vio_stats;
  double *old_state;    // 25 doubles (pose part of old_state)
  double *G;            // 19 x 7 last accepted gain block
  double img_point_cov;
  int level, slot_iter, last_slot;
  int no_publish;       // replicated-solve kernels: every CTA solves, only CTA 0 writes the results to global memory
  unsigned long long *dbg;  // measurement only
};

struct SolveSmem {
  double P[19 * 19];
  double A[49];     // m x m information block
  double HTz[8];
  double vec[19];
  double sol[19];
  double *W;        // literal mode (solve_mode 1) workspace, 19 x 38 doubles
  double *K;        // literal mode, 19 x 19 doubles
};
struct SolveLiteralScratch {
  double W[19 * 38];
  double K[19 * 19];
};

// Exp(v) of include/utils/so3_math.h:44-58 (identity when |v| <= 1e-5)
__device__ inline void so3_exp(const double v[3], double E[9]) {
  double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int i = 0; i < 9; i++) E[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (nrm > 0.00001) {
    double r[3] = {v[0] / nrm, v[1] / nrm, v[2] / nrm};
    double K[9] = {0.0, -r[2], r[1], r[2], 0.0, -r[0], -r[1], r[0], 0.0};
    double s, cc;
    sincos(nrm, &s, &cc);
    const double c = 1.0 - cc;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double kk = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
        E[i * 3 + j] = E[i * 3 + j] + s * K[i * 3 + j] + c * kk;
      }
  }
}
// Log(R) of include/utils/so3_math.h:61-66
__device__ inline void so3_log(const double R[9], double out[3]) {
  double tr = R[0] + R[4] + R[8];
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : acos(0.5 * (tr - 1));
  double K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double f = (fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / sin(theta));
  for (int i = 0; i < 3; i++) out[i] = f * K[i];
}

// vec = state_propagat (-) state  (common_lib.h:194-206), by lane 0 for the rotation part
__device__ inline void boxminus_warp(const double *prop, const double *st, double *vec, int lane) {
  if (lane == 0) {
    double Rd[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rd[i * 3 + j] = st[S_R + 0 * 3 + i] * prop[S_R + 0 * 3 + j] + st[S_R + 1 * 3 + i] * prop[S_R + 1 * 3 + j] +
                                                   st[S_R + 2 * 3 + i] * prop[S_R + 2 * 3 + j];  // b.rot^T * this.rot
    double l[3];
    so3_log(Rd, l);
    vec[0] = l[0], vec[1] = l[1], vec[2] = l[2];
  }
  if (lane >= 3 && lane < 19) {
    // packed offsets of the additive blocks in error-state order: p(3:6) expo(6) v(7:10) bg(10:13) ba(13:16) g(16:19)
    int off = (lane < 6) ? S_P + (lane - 3) : (lane == 6) ? S_EXPO : (lane < 10) ? S_V + (lane - 7) : (lane < 13) ? S_BG + (lane - 10)
              : (lane < 16) ? S_BA + (lane - 13) : S_G + (lane - 16);
    vec[lane] = prop[off] - st[off];
  }
}

// state (+)= sol  (common_lib.h:182-192)
__device__ inline void boxplus_warp(double *st, const double *sol, int lane) {
  if (lane == 0) {
    double E[9], Rn[9];
    so3_exp(sol, E);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rn[i * 3 + j] = st[S_R + i * 3] * E[j] + st[S_R + i * 3 + 1] * E[3 + j] + st[S_R + i * 3 + 2] * E[6 + j];
    for (int i = 0; i < 9; i++) st[S_R + i] = Rn[i];
  }
  if (lane >= 3 && lane < 19) {
    int off = (lane < 6) ? S_P + (lane - 3) : (lane == 6) ? S_EXPO : (lane < 10) ? S_V + (lane - 7) : (lane < 13) ? S_BG + (lane - 10)
              : (lane < 16) ? S_BA + (lane - 13) : S_G + (lane - 16);
    st[off] += sol[lane];
  }
}

// In-place inverse of a 19x19 in shared memory by one warp (Gauss-Jordan, partial pivoting) — literal mode only.
__device__ inline void inverse19_warp(const double *Ain, double *W /*19x38*/, double *out, int lane) {
  for (int idx = lane; idx < 19 * 38; idx += 32) {
    int r = idx / 38, c = idx % 38;
    W[idx] = (c < 19) ? Ain[r * 19 + c] : ((c - 19) == r ? 1.0 : 0.0);
  }
  __syncwarp();
  for (int k = 0; k < 19; k++) {
    int piv = k;
    double best = fabs(W[k * 38 + k]);
    for (int r = k + 1; r < 19; r++) {
      double v = fabs(W[r * 38 + k]);
      if (v > best) best = v, piv = r;
    }
    __syncwarp();
    if (piv != k)
      for (int c = lane; c < 38; c += 32) {
        double t = W[k * 38 + c];
        W[k * 38 + c] = W[piv * 38 + c];
        W[piv * 38 + c] = t;
      }
    __syncwarp();
    double d = W[k * 38 + k];
    double f[19];
    for (int r = 0; r < 19; r++) f[r] = W[r * 38 + k];
    __syncwarp();
    for (int c = lane; c < 38; c += 32) {
      double pk = W[k * 38 + c] / d;
      W[k * 38 + c] = pk;
      for (int r = 0; r < 19; r++)
        if (r != k) W[r * 38 + c] -= f[r] * pk;
    }
    __syncwarp();
  }
  for (int idx = lane; idx < 361; idx += 32) out[idx] = W[(idx / 19) * 38 + 19 + (idx % 19)];
  __syncwarp();
}

// Gain block x = K_1[lane, 0:m] for every lane < 19. PS = P * pscale (pscale = 1 for LIO, 1/img_point_cov for VIO).
template <int m>
__device__ inline void gain_rows(SolveSmem &sm, double pscale, int solve_mode, int lane, double x[m]) {
  if (solve_mode == 1) {
    // literal: K_1 = (H_T_H + (P*pscale)^-1)^-1
    for (int idx = lane; idx < 361; idx += 32) sm.K[idx] = sm.P[idx] * pscale;
    __syncwarp();
    inverse19_warp(sm.K, sm.W, sm.K, lane);
    for (int idx = lane; idx < m * m; idx += 32) sm.K[(idx / m) * 19 + (idx % m)] += sm.A[idx];
    __syncwarp();
    inverse19_warp(sm.K, sm.W, sm.K, lane);
    if (lane < 19)
      for (int j = 0; j < m; j++) x[j] = sm.K[lane * 19 + j];
    return;
  }
  // Warp-cooperative Gauss-Jordan with partial pivoting on the augmented system  M^T [X^T] = [P[:, :m]^T] :
  // lane c < m owns column c of M^T (= row c of M = I + A P_mm), lane m + r owns the right-hand side of state row r
  // (P[r, 0:m] * pscale). After the sweep lane m + r holds K_1[r, 0:m]. m + 19 <= 32 columns, m pivots.
  double col[m];
  if (lane < m) {
#pragma unroll
    for (int i = 0; i < m; i++) {
      double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < m; k++) s += sm.A[lane * m + k] * (sm.P[k * 19 + i] * pscale);
      col[i] = s;
    }
  } else {
    const int r = (lane - m) < 19 ? (lane - m) : 0;
#pragma unroll
    for (int i = 0; i < m; i++) col[i] = sm.P[r * 19 + i] * pscale;
  }
#pragma unroll
  for (int k = 0; k < m; k++) {
    // pivot row: largest |entry| of column k among rows k..m-1, found by the lane that owns column k
    int p = k;
    double best = fabs(col[k]);
#pragma unroll
    for (int i = k + 1; i < m; i++) {
      const double v = fabs(col[i]);
      if (v > best) best = v, p = i;
    }
    p = __shfl_sync(0xffffffffu, p, k);
#pragma unroll
    for (int i = k + 1; i < m; i++) {
      const bool sw = (p == i);
      const double u = col[k], v = col[i];
      col[k] = sw ? v : u;
      col[i] = sw ? u : v;
    }
    const double pv = __shfl_sync(0xffffffffu, col[k], k);
    double f[m];
#pragma unroll
    for (int i = 0; i < m; i++) f[i] = __shfl_sync(0xffffffffu, col[i], k);
    const double vk = col[k] * (1.0 / pv);
    col[k] = vk;
#pragma unroll
    for (int i = 0; i < m; i++)
      if (i != k) col[i] -= f[i] * vk;
  }
#pragma unroll
  for (int i = 0; i < m; i++) x[i] = __shfl_sync(0xffffffffu, col[i], (lane + m) & 31);
}

__device__ inline double warp_norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

#define SOLVE_THREADS 512

// Staging shared by both solve routines: one global round trip brings P, info, the pose parts of state / prior / old_state.
struct SolveIO {
  double info[INFO_N];
  double st[32];   // first 25 doubles of the packed state (R p expo v bg ba g)
  double pr[32];   // same of state_propagat
  double old[32];  // VIO old_state
  double g[19][8]; // gain block G[:, :m]
  int flags[8];
};

// Needs >= 192 threads. Every global load is issued before the first shared-memory store, so the staging costs one
// L2 round trip. __ldcg: the data was written by other SMs of this same grid when called from the persistent kernels.
__device__ __forceinline__ void solve_load(SolveSmem &sm, SolveIO &io, const SolveArgs &a, bool want_old) {
  const int t = threadIdx.x, nt = blockDim.x;
  const double p0 = (t < 361) ? __ldcg(a.state + S_COV + t) : 0.0;
  const double p1 = (t + nt < 361) ? __ldcg(a.state + S_COV + t + nt) : 0.0;
  const double i0 = (t < INFO_N) ? __ldcg(a.info + t) : 0.0;
  const double s0 = (t < 25) ? __ldcg(a.state + t) : 0.0;
  const double r0 = (t < 25) ? a.prop[t] : 0.0;
  const double o0 = (want_old && t < 25) ? __ldcg(a.old_state + t) : 0.0;
  if (t < 361) sm.P[t] = p0;
  if (t + nt < 361) sm.P[t + nt] = p1;
  if (t < INFO_N) io.info[t] = i0;
  if (t < 25) {
    io.st[t] = s0;
    io.pr[t] = r0;
    if (want_old) io.old[t] = o0;
  }
}

// Diagnostics of the iteration just solved (what the reference prints at voxel_map.cpp:404-405). Not needed by the other
// CTAs, so the persistent kernel writes them after publishing the state.
__device__ __forceinline__ void lio_write_stats(const SolveArgs &a, SolveSmem &sm, SolveIO &io) {
  const int tid = threadIdx.x, iterCount = io.flags[3];
  if (a.lio_stats && iterCount < 8) {
    esikf_lio_stats &S = *a.lio_stats;
    for (int t = tid; t < 36; t += blockDim.x) S.HTH[iterCount][t] = sm.A[t];
    for (int t = tid; t < 6; t += blockDim.x) S.HTz[iterCount][t] = sm.HTz[t];
    for (int t = tid; t < 19; t += blockDim.x) S.solution[iterCount][t] = sm.sol[t];
    if (tid == 0) {
      S.iters = iterCount + 1;
      S.effct_feat_num[iterCount] = (int)io.info[INFO_COUNT];
      S.total_residual[iterCount] = io.info[INFO_ABS];
      S.converged[iterCount] = io.flags[0];
    }
  }
}

// One LIO gain solve + state update (src/voxel_map.cpp:462-499) by the calling block. Returns EKF_stop_flg.
// `ctrl` is the loop-control block the routine reads and updates (global memory for the per-iteration kernels, CTA 0's
// shared-memory copy inside the persistent kernel). `resident`: P / poses / info are already staged in sm / io.
__device__ __noinline__ bool lio_solve_block(const SolveArgs &a, SolveSmem &sm, SolveIO &io, Ctrl &ctrl, bool defer_stats, bool resident) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int iterCount = ctrl.iter;
  const int rematch0 = ctrl.rematch_num;
  dbg_stamp(a.dbg, 16);
  if (!resident) solve_load(sm, io, a, false);
  __syncthreads();
  dbg_stamp(a.dbg, 17);
  double x[6], g[6];
  if (tid < 32) {
    for (int idx = lane; idx < 36; idx += 32) sm.A[idx] = io.info[(idx / 6) * 8 + (idx % 6)];  // H^T R^-1 H
    if (lane < 6) sm.HTz[lane] = io.info[lane * 8 + 6];                                          // H^T R^-1 z
    __syncwarp();
    dbg_stamp(a.dbg, 18);
    gain_rows<6>(sm, 1.0, a.solve_mode, lane, x);
    dbg_stamp(a.dbg, 19);
  } else if (tid < 64) {
    boxminus_warp(io.pr, io.st, sm.vec, lane);  // vec = state_propagat (-) state_, concurrently on warp 1 (:470)
  }
  __syncthreads();
  if (tid < 32) {
    // G[lane, 0:6] = K_1[lane, 0:6] * HTH   (voxel_map.cpp:469)
#pragma unroll
    for (int j = 0; j < 6; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += x[k] * sm.A[k * 6 + j];
      g[j] = s;
    }
    // solution = K_1[:, :6] HTz + vec - G[:, :6] vec[:6]   (:471-472)
    if (lane < 19) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) s1 += x[k] * sm.HTz[k], s2 += g[k] * sm.vec[k];
      sm.sol[lane] = s1 + sm.vec[lane] - s2;
#pragma unroll
      for (int j = 0; j < 6; j++) io.g[lane][j] = g[j];
    }
    __syncwarp();
    dbg_stamp(a.dbg, 20);
    boxplus_warp(io.st, sm.sol, lane);  // state_ += solution (:474)
    dbg_stamp(a.dbg, 21);
    if (lane == 0) {
      const bool converged = (warp_norm3(sm.sol) * 57.3 < 0.01) && (warp_norm3(sm.sol + 3) * 100 < 0.015);  // :477
      int rematch = rematch0;
      if (converged || ((rematch == 0) && (iterCount == (a.max_iterations - 2)))) rematch++;  // :482
      const bool stop = (rematch >= 2) || (iterCount == a.max_iterations - 1);                // :485
      io.flags[0] = converged, io.flags[1] = rematch, io.flags[2] = stop;
    }
  }
  __syncthreads();
  const bool stop = io.flags[2] != 0;
  if (!a.no_publish)
    for (int t = tid; t < 25; t += blockDim.x) a.state[t] = io.st[t];
  if (stop && !a.no_publish) {
    // cov = (I - G) cov   (:489-490); G only has its first 6 columns
    for (int t = tid; t < 361; t += blockDim.x) {
      const int r = t / 19, c = t - 19 * r;
      double s = sm.P[t];
#pragma unroll
      for (int j = 0; j < 6; j++) s -= io.g[r][j] * sm.P[j * 19 + c];
      a.state[S_COV + t] = s;
    }
  }
  io.flags[3] = iterCount;
  __syncthreads();
  dbg_stamp(a.dbg, 22);
  if (tid == 0) {
    ctrl.iter = iterCount + 1;
    ctrl.rematch_num = io.flags[1];
    ctrl.stop = stop ? 1 : 0;
  }
  if (!defer_stats) lio_write_stats(a, sm, io);
  return stop;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) lio_solve_kernel(const SolveArgs a) {
  if (a.ctrl->stop) return;
  __shared__ SolveSmem sm;
  __shared__ SolveIO io;
  __shared__ SolveLiteralScratch lit;
  if (threadIdx.x == 0) sm.W = lit.W, sm.K = lit.K;
  __syncthreads();
  lio_solve_block(a, sm, io, *a.ctrl, false, false);
}

__device__ __forceinline__ void vio_write_stats(const SolveArgs &a, SolveSmem &sm, SolveIO &io, const Ctrl &ctrl) {
  const int tid = threadIdx.x, level = a.level, iteration = a.slot_iter;
  const bool accepted = io.flags[0] != 0, ran = io.flags[2] != 0;
  if (ran && a.vio_stats && level < 8) {
    esikf_vio_stats &S = *a.vio_stats;
    if (accepted && iteration < 8) {
      for (int t = tid; t < 49; t += blockDim.x) S.HTH[level][iteration][t] = sm.A[t];
      for (int t = tid; t < 7; t += blockDim.x) S.HTz[level][iteration][t] = sm.HTz[t];
      for (int t = tid; t < 19; t += blockDim.x) S.solution[level][iteration][t] = sm.sol[t];
    }
    if (tid == 0) {
      if (iteration < 8) S.error_trace[level][iteration] = reinterpret_cast<float *>(io.flags)[4];
      S.iters_per_level[level] = iteration + 1;
      if (accepted) S.accepted_per_level[level] += 1;
      S.total_iters += 1;
    }
  }
}

// One VIO accept/rollback + gain solve (src/vio.cpp:1636-1685) by the calling block; on the last slot also the final
// covariance update (:800). Returns EKF_end of the level.
// OVERLAP (opt-in, same arithmetic): warp 1's boxminus runs concurrently with warp 0's accept test and gain elimination;
// the two warps meet at a named barrier right before the solution needs `vec`, instead of a CTA barrier after the boxminus.
template <bool OVERLAP = false>
__device__ __forceinline__ bool vio_solve_block(const SolveArgs &a, SolveSmem &sm, SolveIO &io, Ctrl &ctrl, bool defer_stats, bool resident) {
  const int tid = threadIdx.x, lane = tid & 31;
  const bool level_done_in = (a.slot_iter == 0) ? false : (ctrl.level_done != 0);   // entering a level: EKF_end = false (vio.cpp:1527)
  const float last_error_in = (a.slot_iter == 0) ? FLT_MAX : ctrl.last_error;       // :1528
  const int has_G_in = ctrl.has_G;
  if (level_done_in && !a.last_slot) return true;
  if (!resident) solve_load(sm, io, a, a.slot_iter != 0);
  __syncthreads();
  if (a.slot_iter == 0)
    for (int t = tid; t < 25; t += blockDim.x) io.old[t] = io.st[t];  // old_state = *state at level entry (:1523)
  __syncthreads();
  const int level = a.level, iteration = a.slot_iter;
  // vec = state_propagat (-) state on warp 1 while warp 0 decides accept / rollback and runs the gain solve (:1664)
  if (tid >= 32 && tid < 64 && !level_done_in) {
    boxminus_warp(io.pr, io.st, sm.vec, lane);
    if (OVERLAP) asm volatile("bar.sync 1, 64;" ::: "memory");  // meets warp 0 below
  }
  if (!OVERLAP) __syncthreads();
  if (tid < 32) {
    bool accepted = false, ekf_end = level_done_in;
    float error = 0.f, last_error = last_error_in;
    if (!level_done_in) {
      // error = sum(res^2) / n_meas as float (vio.cpp:1636)
      const double sum_sq = io.info[7 * 8 + 7];
      const int n_meas = (int)io.info[INFO_COUNT];
      error = __fdiv_rn((float)sum_sq, (float)n_meas);
      if (error <= last_error) {  // :1648
        accepted = true;
        if (lane < 25) io.old[lane] = io.st[lane];  // old_state = *state
        last_error = error;
        for (int idx = lane; idx < 49; idx += 32) sm.A[idx] = io.info[(idx / 7) * 8 + (idx % 7)];  // H^T H 7x7
        if (lane < 7) sm.HTz[lane] = io.info[lane * 8 + 7];
        __syncwarp();
        double x[7];
        gain_rows<7>(sm, 1.0 / a.img_point_cov, a.solve_mode, lane, x);
        if (OVERLAP) asm volatile("bar.sync 1, 64;" ::: "memory");  // vec = state_propagat (-) state is complete
        double g[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 7; k++) s += x[k] * sm.A[k * 7 + j];
          g[j] = s;
        }
        if (lane < 19) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int k = 0; k < 7; k++) s1 += x[k] * sm.HTz[k], s2 += g[k] * sm.vec[k];
          sm.sol[lane] = -s1 + sm.vec[lane] - s2;  // :1667
#pragma unroll
          for (int j = 0; j < 7; j++) io.g[lane][j] = g[j];  // G.block<19,7>  (:1665)
        }
        __syncwarp();
        boxplus_warp(io.st, sm.sol, lane);
        // :1675 (float constants 57.3f / 100.0f / 0.001f promote to double against the double norm)
        ekf_end = (warp_norm3(sm.sol) * (double)57.3f < (double)0.001f) && (warp_norm3(sm.sol + 3) * (double)100.0f < (double)0.001f);
      } else {
        if (OVERLAP) asm volatile("bar.sync 1, 64;" ::: "memory");  // warp 1 still reads io.st for the (unused) boxminus
        if (lane < 25) io.st[lane] = io.old[lane];  // *state = old_state  (:1679)
        ekf_end = true;
      }
    }
    if (lane == 0) {
      io.flags[0] = accepted, io.flags[1] = ekf_end, io.flags[2] = !level_done_in;
      reinterpret_cast<float *>(io.flags)[3] = last_error;
      reinterpret_cast<float *>(io.flags)[4] = error;
    }
  }
  __syncthreads();
  const bool accepted = io.flags[0] != 0, ran = io.flags[2] != 0;
  if (ran && !a.no_publish) {
    for (int t = tid; t < 25; t += blockDim.x) {
      a.state[t] = io.st[t];
      a.old_state[t] = io.old[t];
    }
    if (accepted)
      for (int t = tid; t < 133; t += blockDim.x) a.G[t] = io.g[t / 7][t % 7];
    if (!defer_stats) vio_write_stats(a, sm, io, ctrl);
  }
  if (a.last_slot) {
    // state->cov -= G * state->cov   (vio.cpp:800) with the last accepted G (this slot's if accepted, else the stored one)
    const bool haveG = accepted || has_G_in;
    if (haveG)
      for (int t = tid; t < 361; t += blockDim.x) {
        const int r = t / 19, c = t - 19 * r;
        double s = 0.0;
        for (int j = 0; j < 7; j++) s += (accepted ? io.g[r][j] : __ldcg(a.G + r * 7 + j)) * sm.P[j * 19 + c];
        a.state[S_COV + t] = sm.P[t] - s;
      }
  }
  __syncthreads();
  if (tid == 0) {
    if (ran) {
      ctrl.iter += 1;
      ctrl.last_error = reinterpret_cast<float *>(io.flags)[3];
      if (a.slot_iter == 0) ctrl.accepted_in_level = 0;
      if (accepted) ctrl.has_G = 1, ctrl.accepted_in_level += 1;
    }
    ctrl.level_done = io.flags[1];
    if (a.last_slot) ctrl.stop = 1;
  }
  __syncthreads();
  return io.flags[1] != 0;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) vio_solve_kernel(const SolveArgs a) {
  __shared__ SolveSmem sm;
  __shared__ SolveIO io;
  __shared__ SolveLiteralScratch lit;
  if (threadIdx.x == 0) sm.W = lit.W, sm.K = lit.K;
  __syncthreads();
  vio_solve_block(a, sm, io, *a.ctrl, false, false);
}

}  // namespace esikf
