// Persistent update kernels of the B200 ESIKF update (sm_100a): the WHOLE iteration loop of one LIO / VIO update in a
// single cooperative launch, so a tick costs one launch instead of 2 x iterations, the state never leaves the chip between
// iterations and no kernel boundary separates the residual build from the gain solve.
//
//   lio_update_kernel : VoxelMapManager::StateEstimation's loop          (reference src/voxel_map.cpp:372-500)
//   vio_update_kernel : VIOManager::computeJacobianAndUpdateEKF's loops   (src/vio.cpp:784-802, 1520-1688)
//
// Per iteration: every CTA (one per SM, co-resident) builds the residual / Jacobian rows of its slice of the points / patches
// and contracts them on the fp64 tensor-core path; per-CTA 8x8 partial blocks go to global memory; grid barrier; CTA 0 sums
// them in a fixed order, runs the m x m gain solve and the boxplus and publishes the new state; grid barrier; everybody
// reloads the 30 pose / covariance doubles it needs and continues. Results are bit-identical to the per-iteration kernels.
#include "esikf_dev.cuh"

namespace esikf {

// Optional phase timestamps (ns, %globaltimer) written by CTA 0 / thread 0 — measurement only.
__device__ __forceinline__ void stamp(unsigned long long *stamps, int &k) {
  if (stamps && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[k] = t;
  }
  k++;
}

// ---------------------------------------------------------------------------------------------------------------------
// In-kernel all-reduce of the 72-double information buffer over NVLink peer memory (one process per GPU, mailboxes
// exchanged with CUDA IPC), low-latency flavour: no fence, no separate flag. Every rank owns a mailbox
// [2 parities][nranks][72 elements][2 words] in its own HBM. A double travels as two 64-bit words {payload half | 32-bit
// sequence tag}; aligned 64-bit stores are single-copy atomic, so a reader that sees the expected tag in both words has
// the whole value — one NVLink one-way trip instead of store + system fence + flag. CTA 0 of rank r stores its 72
// elements into slot (parity, r) of EVERY rank's mailbox; its threads then spin on the nranks slots of their OWN mailbox
// and add them in rank order, so every rank forms the bit-identical sum with no broadcast, no NCCL call and no kernel
// boundary. Two parities suffice: a rank reaches iteration k+2 only after every peer consumed iteration k.
#define PEER_SLOT_WORDS (INFO_N * 2)  // u64 words per (parity, rank) slot
#define PEER_MAX_RANKS 8
struct PeerArgs {
  unsigned long long *const *mbox;  // device array: mailbox base of every rank (own entry = local memory)
  int rank, nranks;
  unsigned int seq_base;            // tags of this launch are seq_base + iteration + 1 (monotonic, never 0)
};

__device__ __forceinline__ void peer_allreduce(double *info, const PeerArgs &p, unsigned int it) {
  if (p.nranks <= 1) return;
  const int tid = threadIdx.x;
  const unsigned int tag = p.seq_base + it + 1u;
  const int par = it & 1;
  if (tid < INFO_N) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(info[tid]);
    const unsigned long long w0 = (bits << 32) | tag;                       // low half | tag
    const unsigned long long w1 = (bits & 0xffffffff00000000ull) | tag;     // high half | tag
    const size_t off = (size_t)(par * p.nranks + p.rank) * PEER_SLOT_WORDS + 2 * tid;
    for (int r = 0; r < p.nranks; r++) {
      unsigned long long *dst = p.mbox[r] + off;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");
    }
    const unsigned long long *own = p.mbox[p.rank] + (size_t)par * p.nranks * PEER_SLOT_WORDS + 2 * tid;
    double s = 0.0;
    for (int r = 0; r < p.nranks; r++) {
      const unsigned long long *src = own + (size_t)r * PEER_SLOT_WORDS;
      unsigned long long a, b;
      do {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
      } while ((unsigned int)a != tag || (unsigned int)b != tag);
      s += __longlong_as_double((long long)((b & 0xffffffff00000000ull) | (a >> 32)));
    }
    info[tid] = s;
  }
  __syncthreads();
}

// The same exchange split for the replicated-solve kernels: CTA 0 pushes the rank's (locally reduced) buffer into every
// rank's mailbox — its own included —, EVERY CTA pulls the nranks slots of the local mailbox and adds them in rank order.
// `seq` counts executed iterations (tag and parity follow it). Two parities suffice: a rank's CTA 0 writes exchange k + 2
// only after it received every peer's k + 1, which a peer's CTA 0 sends after its own grid barrier k + 1 — and every CTA of
// that peer passes that barrier only after it finished pulling exchange k.
__device__ __forceinline__ void peer_push(const double *info, const PeerArgs &p, unsigned int seq) {
  const int tid = threadIdx.x;
  const unsigned int tag = p.seq_base + seq + 1u;
  const int par = seq & 1;
  if (tid < INFO_N) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(info[tid]);
    const unsigned long long w0 = (bits << 32) | tag;
    const unsigned long long w1 = (bits & 0xffffffff00000000ull) | tag;
    const size_t off = (size_t)(par * p.nranks + p.rank) * PEER_SLOT_WORDS + 2 * tid;
    for (int r = 0; r < p.nranks; r++) {
      unsigned long long *dst = p.mbox[r] + off;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");
    }
  }
}
__device__ __forceinline__ void peer_pull(double *info, const PeerArgs &p, unsigned int seq) {
  const int tid = threadIdx.x;
  const unsigned int tag = p.seq_base + seq + 1u;
  const int par = seq & 1;
  __syncthreads();  // everybody is done reading the local sum (CTA 0's pushers read it too)
  if (tid < INFO_N) {
    const unsigned long long *own = p.mbox[p.rank] + (size_t)par * p.nranks * PEER_SLOT_WORDS + 2 * tid;
    double s = 0.0;
    for (int r = 0; r < p.nranks; r++) {
      const unsigned long long *src = own + (size_t)r * PEER_SLOT_WORDS;
      unsigned long long a, b;
      do {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
      } while ((unsigned int)a != tag || (unsigned int)b != tag);
      s += __longlong_as_double((long long)((b & 0xffffffff00000000ull) | (a >> 32)));
    }
    info[tid] = s;
  }
  __syncthreads();
}

// Counting grid barrier over all CTAs of a cooperative launch: arrivals are atomic increments, everybody polls the same
// counter; the k-th barrier (k = 0, 1, ...) completes when it reaches (k + 1) * gridDim.x. The counter is zeroed for this
// launch by the previous launch (launches alternate between two counters). A variant with a separate release word written
// by the last arriver was measured and is ~1 us SLOWER per barrier (extra L2 hop), see profiles/README.md.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = (epoch + 1) * gridDim.x;
    // release-increment: orders this CTA's earlier writes (made visible to thread 0 by the CTA barrier above) before the
    // arrival, without a separate membar; the acquire poll orders the other CTAs' writes before everything after it.
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  epoch++;
  __syncthreads();
}

// Fixed-order sum of the per-CTA partial blocks (entry-major [entry][block]) into info[] by the calling CTA.
__device__ __forceinline__ void reduce_partials_block(const double *partials, int partial_stride, int nb, double *info) {
  const int tid = threadIdx.x;
  sum_partials(partials, partial_stride, nb, info, blockDim.x >> 5);
  if (tid >= 66 && tid < INFO_N) info[tid] = 0.0;
  __syncthreads();  // `info` is CTA 0's shared-memory copy: no fence needed
}

// Per-CTA partial block -> global (entry-major), same layout / order as reduce_info.
template <int WARPS>
__device__ __forceinline__ void store_partials(ReduceSmem<WARPS> &rs, double D0, double D1, double cnt, bool abs_in_77, double *partials,
                                               int partial_stride) {
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
}

struct FusedSolveSmem {
  SolveSmem sm;
  SolveIO io;
  Ctrl ctrl;  // CTA 0's working copy of the loop-control block (published to global memory after every solve)
};
static_assert(sizeof(FusedSolveSmem) % 8 == 0, "what follows it in shared memory holds doubles");

__global__ void __launch_bounds__(LIO_THREADS, 1) lio_update_kernel(const LioKernelArgs a, const SolveArgs sa, unsigned int *barrier, unsigned int *barrier_next, unsigned long long *stamps,
                                                                     const PeerArgs peer) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LioSmem &sm = *reinterpret_cast<LioSmem *>(smem_raw);
  // CTA 0's solve scratch lives in the reduction scratch (free between the two barriers); only the literal-mode
  // workspace (solve_mode 1, parity checks) borrows the record slots and forces a re-stage.
  static_assert(sizeof(FusedSolveSmem) <= sizeof(sm.fs_raw), "solve scratch must fit");
  FusedSolveSmem &fs = *reinterpret_cast<FusedSolveSmem *>(sm.fs_raw);  // resident for the whole update in CTA 0: P, poses and loop control are staged once
  static_assert(sizeof(SolveLiteralScratch) <= sizeof(sm.rec), "literal scratch must fit in the record slots");
  if (threadIdx.x == 0) {
    SolveLiteralScratch *lit = reinterpret_cast<SolveLiteralScratch *>(&sm.rec[0][0][0]);
    fs.sm.W = lit->W, fs.sm.K = lit->K;
  }
  unsigned int epoch = 0;
  int lo, hi;
  lio_block_range(a.count, lo, hi);
  if (blockIdx.x == 0) {
    // CTA 0 owns the loop control for the whole update: initialise it, the diagnostics and the NEXT launch's barrier
    // counter here (no host-side memset per update). `barrier` alternates between two counters from launch to launch.
    if (threadIdx.x == 0) {
      Ctrl z;
      memset(&z, 0, sizeof(z));
      fs.ctrl = z;
      barrier_next[0] = 0u, barrier_next[32] = 0u;
    }
    if (sa.lio_stats)
      for (int t = threadIdx.x; t < (int)(sizeof(esikf_lio_stats) / 4); t += blockDim.x) reinterpret_cast<int *>(sa.lio_stats)[t] = 0;
    __syncthreads();
  }
  int sk = 0;
  LaneCache lc;  // what stays with this lane's point across iterations (registers + its shared-memory slot)
  lc.staged_idx = -1, lc.have_pt = false, lc.px = lc.py = lc.pz = 0.f;
  for (int it = 0; it < sa.max_iterations; it++) {
    stamp(stamps, sk);  // 0: iteration start
    lio_load_consts(sm, a);
    stamp(stamps, sk);  // 1: constants loaded
    double D0 = 0.0, D1 = 0.0;
    int cnt = 0;
    lio_process_range(a, sm, lo, hi, D0, D1, cnt, lc, it == 0);
    __syncthreads();
    if (stamps && it == 3 && threadIdx.x == 0) {  // measurement only: when did every CTA finish its slice of iteration 3?
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      stamps[640 + blockIdx.x] = t;
    }
    stamp(stamps, sk);  // 2: CTA 0 finished its slice
    store_partials<LIO_WARPS>(sm.red, D0, D1, (double)cnt, true, a.partials, a.partial_stride);
    grid_barrier(barrier, epoch);
    stamp(stamps, sk);  // 3: all CTAs arrived
    if (blockIdx.x == 0) {
      dbg_stamp(a.dbg, 12);
      if (it == 0) {  // stage P / poses / loop control once; later iterations find them in shared memory
        solve_load(fs.sm, fs.io, sa, false);
      }
      reduce_partials_block(a.partials, a.partial_stride, gridDim.x, fs.io.info);
      peer_allreduce(fs.io.info, peer, (unsigned int)it);
      dbg_stamp(a.dbg, 13);
      stamp(stamps, sk);  // 4: partials summed
      if (threadIdx.x == 0) {
        SolveLiteralScratch *lit = reinterpret_cast<SolveLiteralScratch *>(&sm.rec[0][0][0]);
        fs.sm.W = lit->W, fs.sm.K = lit->K;
      }
      __syncthreads();
      lio_solve_block(sa, fs.sm, fs.io, fs.ctrl, true, true);
      if (threadIdx.x == 0) *a.ctrl = fs.ctrl;  // publish stop / iteration count
      if (sa.solve_mode == 1) lc.staged_idx = -1;  // the literal workspace borrowed CTA 0's record slots
    } else {
      sk++;
    }
    stamp(stamps, sk);  // 5: solved
    grid_barrier(barrier, epoch);
    if (blockIdx.x == 0) lio_write_stats(sa, fs.sm, fs.io);
    stamp(stamps, sk);  // 6: state published
    sk++;               // 7: spare
    if (__ldcg(&a.ctrl->stop)) break;  // EKF_stop_flg (voxel_map.cpp:499)
  }
}

__global__ void __launch_bounds__(VIO_THREADS, 1) vio_update_kernel(const VioKernelArgs a, SolveArgs sa, unsigned int *barrier, unsigned int *barrier_next, unsigned long long *stamps,
                                                                     const PeerArgs peer) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  VioSmem &sm = *reinterpret_cast<VioSmem *>(smem_raw);
  // the solve scratch has its own shared memory behind VioSmem (so the diagnostics can be written while the next
  // iteration already stages rows); only the literal-mode workspace borrows the row area
  FusedSolveSmem &fs = *reinterpret_cast<FusedSolveSmem *>(smem_raw + sizeof(VioSmem));
  struct VioFused {
    SolveLiteralScratch lit;
  };
  VioFused &vf = *reinterpret_cast<VioFused *>(&sm.rows[0][0][0]);
  static_assert(sizeof(VioFused) <= sizeof(sm.rows), "literal scratch must fit in the row staging area");
  unsigned int epoch = 0;
  int lo, hi;
  vio_block_range(a.count, lo, hi);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      Ctrl z;
      memset(&z, 0, sizeof(z));
      fs.ctrl = z;
      barrier_next[0] = 0u, barrier_next[32] = 0u;
    }
    if (sa.vio_stats)
      for (int t = threadIdx.x; t < (int)(sizeof(esikf_vio_stats) / 4); t += blockDim.x) reinterpret_cast<int *>(sa.vio_stats)[t] = 0;
    __syncthreads();
  }
  for (int level = a.levels - 1; level >= 0; level--) {      // vio.cpp:790
    for (int it = 0; it < sa.max_iterations; it++) {          // :1536
      int sk = 8 * ((a.levels - 1 - level) * sa.max_iterations + it);
      stamp(stamps, sk);
      vio_load_consts(sm, a);
      stamp(stamps, sk);
      double D0 = 0.0, D1 = 0.0, n_meas = 0.0;
      vio_process_range(a, sm, level, lo, hi, D0, D1, n_meas);
      __syncthreads();
      stamp(stamps, sk);
      store_partials<VIO_WARPS>(sm.red, D0, D1, n_meas, false, a.partials, a.partial_stride);
      grid_barrier(barrier, epoch);
      stamp(stamps, sk);
      bool level_done;
      if (blockIdx.x == 0) {
        if (level == a.levels - 1 && it == 0) {
          solve_load(fs.sm, fs.io, sa, false);
        }
        reduce_partials_block(a.partials, a.partial_stride, gridDim.x, fs.io.info);
        peer_allreduce(fs.io.info, peer, (unsigned int)((a.levels - 1 - level) * sa.max_iterations + it));
        stamp(stamps, sk);
        sa.level = level, sa.slot_iter = it, sa.last_slot = 0;
        if (threadIdx.x == 0) fs.sm.W = vf.lit.W, fs.sm.K = vf.lit.K;
        __syncthreads();
        vio_solve_block(sa, fs.sm, fs.io, fs.ctrl, true, true);
        if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
      } else {
        sk++;
      }
      stamp(stamps, sk);
      grid_barrier(barrier, epoch);
      if (blockIdx.x == 0) vio_write_stats(sa, fs.sm, fs.io, fs.ctrl);
      stamp(stamps, sk);
      level_done = __ldcg(&a.ctrl->level_done) != 0;
      if (level_done) break;  // EKF_end (:1685)
    }
  }
  // state->cov -= G * state->cov (vio.cpp:800): a last-slot pass of the solve routine with the level already finished
  if (blockIdx.x == 0) {
    __syncthreads();
    sa.level = 0, sa.slot_iter = 1, sa.last_slot = 1;
    if (threadIdx.x == 0) {
      fs.ctrl.level_done = 1;
      fs.sm.W = vf.lit.W, fs.sm.K = vf.lit.K;
    }
    __syncthreads();
    vio_solve_block(sa, fs.sm, fs.io, fs.ctrl, false, true);
    if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Replicated-solve variants (loop_mode 2, the default on a single GPU). Every CTA keeps its own copy of P, the poses and the loop control
// in shared memory, sums the per-CTA partials itself after the one grid barrier of the iteration and runs the same m x m
// solve — same instructions on the same inputs in the same order, so all copies stay bit-identical and equal to what
// CTA 0 computes in the kernels above. That removes the second grid barrier, the publication of the state through global
// memory and the reload of the pose / covariance blocks from every iteration; only CTA 0 writes results and diagnostics.
// The partial blocks are double-buffered by iteration parity: a CTA can overwrite buffer p again only two iterations
// later, after a barrier that every reader of iteration k has already passed. Measured on config 2: 61.4 k it/s against
// 55.1 k for the CTA-0 solve (profiles/loop_modes_r01_mode2.txt). Exchanging the partials as tagged 64-bit words polled
// by every CTA instead of the barrier was also built and measured: bit-identical but slower (49.3 k it/s, the polling
// of 148 CTAs saturates L2; profiles/loop_modes_r01_mode3_ll.txt) and removed again.
// grid_barrier whose waiting time does useful work: thread 0 arrives and polls as in grid_barrier, the other threads run
// `work` (global writes nobody reads inside the kernel) meanwhile. Same counter protocol, so both forms can be mixed.
template <class F> __device__ __forceinline__ void grid_barrier_overlap(unsigned int *counter, unsigned int &epoch, F work) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = (epoch + 1) * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  } else {
    work();
  }
  epoch++;
  __syncthreads();
}

// Diagnostics of the last solved iteration written by ONE warp (lane-strided) — the deferred form of lio_write_stats /
// vio_write_stats + the control-block publication, run by warp 1 of CTA 0 inside the next barrier wait.
__device__ __forceinline__ void lio_write_stats_warp(const SolveArgs &a, const SolveSmem &sm, const SolveIO &io, const Ctrl &ctrl, Ctrl *ctrl_out, int lane) {
  const int iterCount = io.flags[3];
  if (lane < (int)(sizeof(Ctrl) / 4)) reinterpret_cast<int *>(ctrl_out)[lane] = reinterpret_cast<const int *>(&ctrl)[lane];
  if (a.lio_stats && iterCount < 8) {
    esikf_lio_stats &S = *a.lio_stats;
    for (int t = lane; t < 36; t += 32) S.HTH[iterCount][t] = sm.A[t];
    if (lane < 6) S.HTz[iterCount][lane] = sm.HTz[lane];
    if (lane < 19) S.solution[iterCount][lane] = sm.sol[lane];
    if (lane == 31) {
      S.iters = iterCount + 1;
      S.effct_feat_num[iterCount] = (int)io.info[INFO_COUNT];
      S.total_residual[iterCount] = io.info[INFO_ABS];
      S.converged[iterCount] = io.flags[0];
    }
  }
}
__device__ __forceinline__ void vio_write_stats_warp(const SolveArgs &a, int level, int iteration, const SolveSmem &sm, const SolveIO &io, const Ctrl &ctrl,
                                                     Ctrl *ctrl_out, int lane) {
  const bool accepted = io.flags[0] != 0, ran = io.flags[2] != 0;
  if (lane < (int)(sizeof(Ctrl) / 4)) reinterpret_cast<int *>(ctrl_out)[lane] = reinterpret_cast<const int *>(&ctrl)[lane];
  if (ran && a.vio_stats && level < 8) {
    esikf_vio_stats &S = *a.vio_stats;
    if (accepted && iteration < 8) {
      for (int t = lane; t < 49; t += 32) S.HTH[level][iteration][t] = sm.A[t];
      if (lane < 7) S.HTz[level][iteration][lane] = sm.HTz[lane];
      if (lane < 19) S.solution[level][iteration][lane] = sm.sol[lane];
    }
    if (lane == 31) {
      if (iteration < 8) S.error_trace[level][iteration] = reinterpret_cast<const float *>(io.flags)[4];
      S.iters_per_level[level] = iteration + 1;
      if (accepted) S.accepted_per_level[level] += 1;
      S.total_iters += 1;
    }
  }
}

__device__ __forceinline__ void lio_consts_from_resident(LioSmem &sm, const FusedSolveSmem &fs) {
  const int tid = threadIdx.x;
  if (tid < 9) {
    const int r = tid / 3, c = tid % 3;
    sm.R[tid] = fs.io.st[S_R + tid];
    sm.Ptt[tid] = fs.sm.P[r * 19 + c];
    sm.Ppp[tid] = fs.sm.P[(3 + r) * 19 + (3 + c)];
  } else if (tid < 12) {
    sm.t[tid - 9] = fs.io.st[S_P + tid - 9];
  }
  __syncthreads();
}

// DEAL: 32-point chunks dealt round-robin over the CTAs instead of one contiguous block per CTA (see lio_process_range).
// DEFER: CTA 0 writes the diagnostics / control block of iteration k while it waits at the grid barrier of iteration k + 1
// (warp 1, while thread 0 polls) instead of right after the solve, where it delays CTA 0's next slice — and with it the
// whole grid — by the ~1 us the "publish" phase takes in profiles/loop_modes_r01_mode2.txt.
// PEER: every CTA pulls the peer-reduced information buffer from the local NVLink mailbox (CTA 0 pushed it), see peer_push.
template <bool DEAL, bool DEFER, bool PEER>
__device__ __forceinline__ void lio_update_repl_body(const LioKernelArgs &a, const SolveArgs &sa_in, unsigned int *barrier, unsigned int *barrier_next,
                                                     unsigned long long *stamps, size_t partial_parity_stride, const PeerArgs &peer) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LioSmem &sm = *reinterpret_cast<LioSmem *>(smem_raw);
  FusedSolveSmem &fs = *reinterpret_cast<FusedSolveSmem *>(sm.fs_raw);
  SolveArgs sa = sa_in;
  sa.no_publish = (blockIdx.x != 0);
  if (blockIdx.x != 0) sa.dbg = nullptr;
  unsigned int epoch = 0;
  int lo, hi;
  if (DEAL)
    lo = 0, hi = a.count;
  else
    lio_block_range(a.count, lo, hi);
  if (threadIdx.x == 0) {
    Ctrl z;
    memset(&z, 0, sizeof(z));
    fs.ctrl = z;
    if (blockIdx.x == 0) barrier_next[0] = 0u, barrier_next[32] = 0u;
  }
  if (blockIdx.x == 0 && sa.lio_stats)
    for (int t = threadIdx.x; t < (int)(sizeof(esikf_lio_stats) / 4); t += blockDim.x) reinterpret_cast<int *>(sa.lio_stats)[t] = 0;
  solve_load(fs.sm, fs.io, sa, false);  // P, current and prior pose: staged once per CTA, resident for the whole update
  {
    // prior-pose constants (voxel_map.cpp:425-428, 445) never change inside the loop
    const int tid = threadIdx.x;
    if (tid < 9) {
      const int r = tid / 3, c = tid % 3;
      sm.Rp[tid] = a.prop[S_R + tid];
      double s = 0;
      for (int k = 0; k < 3; k++) s += a.prop[S_R + r * 3 + k] * a.extR[k * 3 + c];
      sm.Mp[tid] = s;
    } else if (tid < 12) {
      sm.tp[tid - 9] = a.prop[S_P + tid - 9];
    }
  }
  __syncthreads();
  int sk = 0;
  LaneCache lc;
  lc.staged_idx = -1, lc.have_pt = false, lc.px = lc.py = lc.pz = 0.f;
  for (int it = 0; it < sa.max_iterations; it++) {
    stamp(stamps, sk);  // 0: iteration start
    lio_consts_from_resident(sm, fs);
    stamp(stamps, sk);  // 1: constants in place
    double D0 = 0.0, D1 = 0.0;
    int cnt = 0;
    lio_process_range<DEAL>(a, sm, lo, hi, D0, D1, cnt, lc, it == 0);
    __syncthreads();
    if (stamps && it == 3 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      stamps[640 + blockIdx.x] = t;
    }
    stamp(stamps, sk);  // 2: CTA 0 finished its slice
    double *const part = a.partials + (size_t)(it & 1) * partial_parity_stride;
    store_partials<LIO_WARPS>(sm.red, D0, D1, (double)cnt, true, part, a.partial_stride);
    if (DEFER && blockIdx.x == 0 && it > 0)
      grid_barrier_overlap(barrier, epoch, [&]() {
        if ((threadIdx.x >> 5) == 1) lio_write_stats_warp(sa, fs.sm, fs.io, fs.ctrl, a.ctrl, threadIdx.x & 31);
      });
    else
      grid_barrier(barrier, epoch);
    stamp(stamps, sk);  // 3: all CTAs arrived
    reduce_partials_block(part, a.partial_stride, gridDim.x, fs.io.info);
    if (PEER && peer.nranks > 1) {
      if (blockIdx.x == 0) peer_push(fs.io.info, peer, (unsigned int)it);
      peer_pull(fs.io.info, peer, (unsigned int)it);
    }
    stamp(stamps, sk);  // 4: partials summed
    if (threadIdx.x == 0) {
      SolveLiteralScratch *lit = reinterpret_cast<SolveLiteralScratch *>(&sm.rec[0][0][0]);
      fs.sm.W = lit->W, fs.sm.K = lit->K;
    }
    __syncthreads();
    lio_solve_block(sa, fs.sm, fs.io, fs.ctrl, true, true);
    if (sa.solve_mode == 1) lc.staged_idx = -1;  // the literal workspace borrowed the record slots
    stamp(stamps, sk);  // 5: solved
    __syncthreads();    // fs.ctrl (written by thread 0) is read by everybody below
    const bool last = fs.ctrl.stop || it == sa.max_iterations - 1;
    if (blockIdx.x == 0 && (!DEFER || last)) {
      if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
      lio_write_stats(sa, fs.sm, fs.io);
    }
    stamp(stamps, sk);  // 6
    sk++;               // 7: spare
    if (fs.ctrl.stop) break;  // EKF_stop_flg (voxel_map.cpp:499)
  }
}

template <bool DEAL, bool DEFER>
__global__ void __launch_bounds__(LIO_THREADS, 1) lio_update_repl_kernel(const LioKernelArgs a, const SolveArgs sa_in, unsigned int *barrier, unsigned int *barrier_next,
                                                                          unsigned long long *stamps, size_t partial_parity_stride) {
  PeerArgs none;
  none.mbox = nullptr, none.rank = 0, none.nranks = 1, none.seq_base = 0;
  lio_update_repl_body<DEAL, DEFER, false>(a, sa_in, barrier, barrier_next, stamps, partial_parity_stride, none);
}
// the same loop with peer GPUs attached (ESIKF_TUNE_PEER_REPLICATED)
__global__ void __launch_bounds__(LIO_THREADS, 1) lio_update_repl_peer_kernel(const LioKernelArgs a, const SolveArgs sa_in, unsigned int *barrier, unsigned int *barrier_next,
                                                                               unsigned long long *stamps, size_t partial_parity_stride, const PeerArgs peer) {
  lio_update_repl_body<false, false, true>(a, sa_in, barrier, barrier_next, stamps, partial_parity_stride, peer);
}

__device__ __forceinline__ void vio_consts_from_resident(VioSmem &sm, const VioKernelArgs &a, const FusedSolveSmem &fs) {
  const int tid = threadIdx.x;
  if (tid < 9) {
    // Rcw = Rci * Rwi^T  (vio.cpp:1542)
    const int r = tid / 3, c = tid % 3;
    double s = 0;
    for (int k = 0; k < 3; k++) s += a.Rci[r * 3 + k] * fs.io.st[S_R + c * 3 + k];
    sm.Rcw[tid] = s;
  }
  if (tid == 0) sm.inv_expo = fs.io.st[S_EXPO];
  __syncthreads();
  if (tid < 3) {
    // Pcw = -Rci Rwi^T Pwi + Pci  (:1543)
    double s = 0;
    for (int k = 0; k < 3; k++) s += sm.Rcw[tid * 3 + k] * fs.io.st[S_P + k];
    sm.Pcw[tid] = -s + a.Pci[tid];
  }
  __syncthreads();
}

// FAST: per-patch inputs cached across iterations + exact-reciprocal tap-stride arithmetic (vio_process_range<true>) and
// the boxminus overlapped with the gain elimination (vio_solve_block<true>); bit-identical results.
template <bool DEFER, bool FAST, bool PEER>
__device__ __forceinline__ void vio_update_repl_body(const VioKernelArgs &a, SolveArgs sa, unsigned int *barrier, unsigned int *barrier_next,
                                                     unsigned long long *stamps, size_t partial_parity_stride, const PeerArgs &peer) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  VioSmem &sm = *reinterpret_cast<VioSmem *>(smem_raw);
  FusedSolveSmem &fs = *reinterpret_cast<FusedSolveSmem *>(smem_raw + sizeof(VioSmem));
  VioPatchSlot *const slots = reinterpret_cast<VioPatchSlot *>(smem_raw + sizeof(VioSmem) + sizeof(FusedSolveSmem));  // FAST launches only
  VioLaneCache lc;
  lc.Pv = make_float2(0.f, 0.f), lc.level = -1, lc.have = false;
  SolveLiteralScratch &lit = *reinterpret_cast<SolveLiteralScratch *>(&sm.rows[0][0][0]);
  static_assert(sizeof(SolveLiteralScratch) <= sizeof(sm.rows), "literal scratch must fit in the row staging area");
  sa.no_publish = (blockIdx.x != 0);
  if (blockIdx.x != 0) sa.dbg = nullptr;
  unsigned int epoch = 0;
  int lo, hi;
  vio_block_range(a.count, lo, hi);
  if (threadIdx.x == 0) {
    Ctrl z;
    memset(&z, 0, sizeof(z));
    fs.ctrl = z;
    if (blockIdx.x == 0) barrier_next[0] = 0u, barrier_next[32] = 0u;
  }
  if (blockIdx.x == 0 && sa.vio_stats)
    for (int t = threadIdx.x; t < (int)(sizeof(esikf_vio_stats) / 4); t += blockDim.x) reinterpret_cast<int *>(sa.vio_stats)[t] = 0;
  solve_load(fs.sm, fs.io, sa, false);
  __syncthreads();
  int slot = 0;
  int pend_level = -1, pend_it = 0;  // DEFER: iteration whose diagnostics CTA 0 still has to write
  for (int level = a.levels - 1; level >= 0; level--) {      // vio.cpp:790
    for (int it = 0; it < sa.max_iterations; it++) {          // :1536
      const int cur = slot++;  // counts executed iterations (a level may end early): the partial-buffer parity follows it
      int sk = 8 * ((a.levels - 1 - level) * sa.max_iterations + it);
      stamp(stamps, sk);
      vio_consts_from_resident(sm, a, fs);
      stamp(stamps, sk);
      double D0 = 0.0, D1 = 0.0, n_meas = 0.0;
      if (FAST)
        vio_process_range<true>(a, sm, level, lo, hi, D0, D1, n_meas, slots, &lc);
      else
        vio_process_range(a, sm, level, lo, hi, D0, D1, n_meas);
      __syncthreads();
      stamp(stamps, sk);
      double *const part = a.partials + (size_t)(cur & 1) * partial_parity_stride;
      store_partials<VIO_WARPS>(sm.red, D0, D1, n_meas, false, part, a.partial_stride);
      if (DEFER && blockIdx.x == 0 && pend_level >= 0) {
        grid_barrier_overlap(barrier, epoch, [&]() {
          if ((threadIdx.x >> 5) == 1) vio_write_stats_warp(sa, pend_level, pend_it, fs.sm, fs.io, fs.ctrl, a.ctrl, threadIdx.x & 31);
        });
        pend_level = -1;
      } else {
        grid_barrier(barrier, epoch);
      }
      stamp(stamps, sk);
      reduce_partials_block(part, a.partial_stride, gridDim.x, fs.io.info);
      if (PEER && peer.nranks > 1) {
        if (blockIdx.x == 0) peer_push(fs.io.info, peer, (unsigned int)cur);
        peer_pull(fs.io.info, peer, (unsigned int)cur);
      }
      stamp(stamps, sk);
      sa.level = level, sa.slot_iter = it, sa.last_slot = 0;
      if (threadIdx.x == 0) fs.sm.W = lit.W, fs.sm.K = lit.K;
      __syncthreads();
      vio_solve_block<FAST>(sa, fs.sm, fs.io, fs.ctrl, true, true);  // ends with a CTA barrier: fs.ctrl is current for everybody
      stamp(stamps, sk);
      if (blockIdx.x == 0) {
        if (DEFER) {
          pend_level = level, pend_it = it;
        } else {
          if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
          vio_write_stats(sa, fs.sm, fs.io, fs.ctrl);
        }
      }
      stamp(stamps, sk);
      if (fs.ctrl.level_done) break;  // EKF_end (:1685)
    }
  }
  if (DEFER && blockIdx.x == 0 && pend_level >= 0) {  // the last iteration's diagnostics: nothing left to hide them behind
    sa.level = pend_level, sa.slot_iter = pend_it;
    if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
    vio_write_stats(sa, fs.sm, fs.io, fs.ctrl);
  }
  // state->cov -= G * state->cov (vio.cpp:800): a last-slot pass of the solve routine with the level already finished
  if (blockIdx.x == 0) {
    __syncthreads();
    sa.level = 0, sa.slot_iter = 1, sa.last_slot = 1;
    if (threadIdx.x == 0) {
      fs.ctrl.level_done = 1;
      fs.sm.W = lit.W, fs.sm.K = lit.K;
    }
    __syncthreads();
    vio_solve_block(sa, fs.sm, fs.io, fs.ctrl, false, true);
    if (threadIdx.x == 0) *a.ctrl = fs.ctrl;
  }
}

template <bool DEFER, bool FAST>
__global__ void __launch_bounds__(VIO_THREADS, 1) vio_update_repl_kernel(const VioKernelArgs a, SolveArgs sa, unsigned int *barrier, unsigned int *barrier_next,
                                                                          unsigned long long *stamps, size_t partial_parity_stride) {
  PeerArgs none;
  none.mbox = nullptr, none.rank = 0, none.nranks = 1, none.seq_base = 0;
  vio_update_repl_body<DEFER, FAST, false>(a, sa, barrier, barrier_next, stamps, partial_parity_stride, none);
}
__global__ void __launch_bounds__(VIO_THREADS, 1) vio_update_repl_peer_kernel(const VioKernelArgs a, SolveArgs sa, unsigned int *barrier, unsigned int *barrier_next,
                                                                               unsigned long long *stamps, size_t partial_parity_stride, const PeerArgs peer) {
  vio_update_repl_body<false, false, true>(a, sa, barrier, barrier_next, stamps, partial_parity_stride, peer);
}

}  // namespace esikf
