// LIO kernels of the B200 ESIKF update (sm_100a).
//
//   lio_precompute_kernel : per-frame calcBodyCov + cross-matrix vector     (reference src/voxel_map.cpp:15-34, 349-360)
//   lio_residual_kernel   : one ESIKF iteration's residual / Jacobian build (src/voxel_map.cpp:376-390 TransformLidar + point
//                           covariance, :643-786 voxel probe + plane association, :414-458 Jacobian / R^-1) fused with the
//                           H^T R^-1 H, H^T R^-1 z reduction (:464-466). No PointToPlane is ever materialised.
//
// Mapping: one thread per LiDAR point (the ~600 fp64 operations per point are sequential; a warp per point would idle 31/32
// of the fp64 pipe), 22 warps per CTA, one CTA per SM (148 x 704 = 104 k points in one round). The warp is the cooperation
// unit: (i) the 32 first-candidate plane records of a warp are staged into shared memory with coalesced half-warp copies
// and stay resident across the iterations of the persistent kernel; (ii) the rare extra candidates of sub-divided root
// voxels of ALL lanes are evaluated lane-parallel in one pass; (iii) the per-warp contraction sum_i a_i (w_i a_i)^T,
// a = [H_i(6), z_i, 1], runs on the fp64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA) out of the shared-memory rows.
// Partial 8x8 blocks are combined in a fixed order (warp -> CTA -> grid), so results are bit-reproducible run to run.
#include "esikf_dev.cuh"

namespace esikf {

#define LIO_THREADS 704  // 22 warps, one CTA per SM: 148 x 704 = 104k points in a single round
#define LIO_WARPS (LIO_THREADS / 32)

struct LioKernelArgs {
  const float *pts;          // [n_total][3] body-frame scan
  const double *pre;         // SoA [9][pre_stride]: cross vector c(3) | body cov xx xy xz yy yz zz
  int pre_stride;
  int partial_stride;
  int begin, count;          // this rank's shard
  const double *state;       // current iterate (device, packed)
  const double *prop;        // state_propagat
  const HashSlot *slots;
  uint32_t hash_mask;
  const esikf_plane *planes;
  double extR[9], extT[3];
  double voxel_size;         // double voxel size used for the key (voxel_map.cpp:646,668)
  double inv_voxel_size;     // 1 / voxel_size, used when exact
  int inv_voxel_exact;
  float voxel_size_f;        // float voxel size that positioned the roots (voxel_map.cpp:534,578-581)
  double sigma_num;
  int32_t *match_plane;      // [n_total]
  int32_t *normal_plane;     // [n_total] sticky
  float *dis_to_plane;       // [n_total]
  double *partials;          // [grid][INFO_N]
  double *info;              // [INFO_N]
  Ctrl *ctrl;
  unsigned long long *dbg;   // measurement only
  int init_normal;           // first iteration of an update: unmatched points get normal_plane = -1 (pv.normal = 0)
};

__device__ __forceinline__ double dot3_rn(double a0, double a1, double a2, double b0, double b1, double b2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

// index of (i,j) in the row-major upper triangle of a 6x6
__device__ __forceinline__ constexpr int tri6(int i, int j) {
  return (i <= j) ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j));
}

// sigma = J^T PV J evaluated as (J PV) J, the order of src/voxel_map.cpp:735. pv points at the packed upper triangle inside
// the plane record (shared or global memory); the per-column fences keep the 21 loads from being hoisted into one
// 44-register burst (the kernel runs at 80 registers / thread).
__device__ __forceinline__ double quad6(const double *pv, const double J[6]) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double t = J[0] * pv[tri6(0, j)];
#pragma unroll
    for (int i = 1; i < 6; i++) t += J[i] * pv[tri6(i, j)];
    s = (j == 0) ? t * J[0] : s + t * J[j];
    asm volatile("" ::: "memory");
  }
  return s;
}

__device__ __forceinline__ double quad3_sym(const double v[6], double n0, double n1, double n2) {
  // n^T V n with V symmetric (xx xy xz yy yz zz), evaluated as (n^T V) n
  double t0 = n0 * v[0] + n1 * v[1] + n2 * v[2];
  double t1 = n0 * v[1] + n1 * v[3] + n2 * v[4];
  double t2 = n0 * v[2] + n1 * v[4] + n2 * v[5];
  return t0 * n0 + t1 * n1 + t2 * n2;
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-frame precompute: calcBodyCov (voxel_map.cpp:15-34) and the cross-matrix vector extR*p+extT (:356-359).
__global__ void lio_precompute_kernel(const float *__restrict__ pts, int n, double *__restrict__ pre, int pre_stride,
                                      const double *__restrict__ ext, float dept_err, float beam_err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
  if (pz == 0) pz = 0.001;  // :352  (calcBodyCov's own 0 -> 1e-4 fix at :17 can then never trigger)
  float range = (float)sqrt(px * px + py * py + pz * pz);
  float range_var = dept_err * dept_err;
  double sdv = sin((double)beam_err * 0.017453293);  // PCL DEG2RAD
  double dv = sdv * sdv;
  double nrm = sqrt(px * px + py * py + pz * pz);
  double dx = px / nrm, dy = py / nrm, dz = pz / nrm;
  double b1x = 1.0, b1y = 1.0, b1z = -(dx + dy) / dz;
  double n1 = sqrt(b1x * b1x + b1y * b1y + b1z * b1z);
  b1x /= n1, b1y /= n1, b1z /= n1;
  double b2x = b1y * dz - b1z * dy, b2y = b1z * dx - b1x * dz, b2z = b1x * dy - b1y * dx;  // base_vector1.cross(direction)
  double n2 = sqrt(b2x * b2x + b2y * b2y + b2z * b2z);
  b2x /= n2, b2y /= n2, b2z /= n2;
  // A = range * [d]x * [b1 b2]   (3x2)
  double r = (double)range;
  double a00 = r * (-dz * b1y + dy * b1z), a01 = r * (-dz * b2y + dy * b2z);
  double a10 = r * (dz * b1x - dx * b1z), a11 = r * (dz * b2x - dx * b2z);
  double a20 = r * (-dy * b1x + dx * b1y), a21 = r * (-dy * b2x + dx * b2y);
  double rv = (double)range_var;
  double *o = pre + i;
  const size_t ns = (size_t)pre_stride;
  // cross vector
  o[0 * ns] = ext[0] * px + ext[1] * py + ext[2] * pz + ext[9];
  o[1 * ns] = ext[3] * px + ext[4] * py + ext[5] * pz + ext[10];
  o[2 * ns] = ext[6] * px + ext[7] * py + ext[8] * pz + ext[11];
  // cov = d rv d^T + A dv A^T  (symmetric; upper triangle stored)
  o[3 * ns] = dx * rv * dx + dv * (a00 * a00 + a01 * a01);
  o[4 * ns] = dx * rv * dy + dv * (a00 * a10 + a01 * a11);
  o[5 * ns] = dx * rv * dz + dv * (a00 * a20 + a01 * a21);
  o[6 * ns] = dy * rv * dy + dv * (a10 * a10 + a11 * a11);
  o[7 * ns] = dy * rv * dz + dv * (a10 * a20 + a11 * a21);
  o[8 * ns] = dz * rv * dz + dv * (a20 * a20 + a21 * a21);
}

// ---------------------------------------------------------------------------------------------------------------------
struct Cand {
  double prob;
  int idx;
  float dis;  // signed n.p + d narrowed to float (PointToPlane::dis_to_plane_, voxel_map.cpp:753)
};
struct EvalOut {
  bool pass;
  double prob;
  float dis;
};

// build_single_residual's plane branch (src/voxel_map.cpp:721-768) for one candidate plane record `q` (shared or global
// memory, 32 doubles). The probability (:740) is only needed to arbitrate between several candidates.
// n^T pv.var n without forming pv.var (voxel_map.cpp:385-388, 736):
//   pv.var = R body_cov R^T + [c]x P_tt [c]x^T + P_pp  =>  n^T var n = m^T body_cov m + u^T P_tt u + n^T P_pp n,  m = R^T n, u = c x n.
struct PointCovRef {
  const double *pre;  // SoA element of this point: c(3) | body cov (6), stride `ns`
  size_t ns;
  const double *R, *Ptt, *Ppp;  // shared-memory copies of the current rotation / covariance blocks
};
__device__ __forceinline__ double n_var_n(const PointCovRef &pc, double n0, double n1, double n2) {
  const double *R = pc.R;
  const double m0 = R[0] * n0 + R[3] * n1 + R[6] * n2, m1 = R[1] * n0 + R[4] * n1 + R[7] * n2, m2 = R[2] * n0 + R[5] * n1 + R[8] * n2;
  const double *pre = pc.pre;
  const size_t ns = pc.ns;
  const double bc[6] = {pre[3 * ns], pre[4 * ns], pre[5 * ns], pre[6 * ns], pre[7 * ns], pre[8 * ns]};
  double s = quad3_sym(bc, m0, m1, m2);
  const double cx = pre[0], cy = pre[ns], cz = pre[2 * ns];
  const double u0 = cy * n2 - cz * n1, u1 = cz * n0 - cx * n2, u2 = cx * n1 - cy * n0;
  const double *P = pc.Ptt;
  s += (u0 * P[0] + u1 * P[3] + u2 * P[6]) * u0 + (u0 * P[1] + u1 * P[4] + u2 * P[7]) * u1 + (u0 * P[2] + u1 * P[5] + u2 * P[8]) * u2;
  const double *Q = pc.Ppp;
  s += (n0 * Q[0] + n1 * Q[3] + n2 * Q[6]) * n0 + (n0 * Q[1] + n1 * Q[4] + n2 * Q[7]) * n1 + (n0 * Q[2] + n1 * Q[5] + n2 * Q[8]) * n2;
  return s;
}

__device__ __forceinline__ EvalOut eval_rec(const double *__restrict__ q, const double pw[3], const PointCovRef &pc, double sigma_num,
                                            bool need_prob) {
  EvalOut o;
  o.pass = false, o.prob = 0.0, o.dis = 0.f;
  const double2 *__restrict__ q2 = reinterpret_cast<const double2 *>(q);
  const double2 a0 = q2[0], a1 = q2[1], a2 = q2[2];  // c0 c1 | c2 n0 | n1 n2
  const double c0 = a0.x, c1 = a0.y, c2 = a1.x, n0 = a1.y, n1 = a2.x, n2 = a2.y;
  const float2 dr = *reinterpret_cast<const float2 *>(q + 27);  // d, radius
  // float-rounded quantities that gate the association: evaluated without FMA contraction, left to right, like the oracle
  const double sd = __dadd_rn(dot3_rn(n0, n1, n2, pw[0], pw[1], pw[2]), (double)dr.x);
  const float dis_to_plane = (float)fabs(sd);
  const double e0 = c0 - pw[0], e1 = c1 - pw[1], e2 = c2 - pw[2];
  const float dis_to_center = (float)dot3_rn(e0, e1, e2, e0, e1, e2);
  const float range_dis = sqrtf(__fsub_rn(dis_to_center, __fmul_rn(dis_to_plane, dis_to_plane)));
  if ((double)range_dis <= 3.0 * (double)dr.y) {  // NaN fails, as in the reference
    const double J[6] = {pw[0] - c0, pw[1] - c1, pw[2] - c2, -n0, -n1, -n2};
    double sigma_l = quad6(q + 6, J);
    sigma_l += n_var_n(pc, n0, n1, n2);
    if ((double)dis_to_plane < sigma_num * sqrt(sigma_l)) {
      o.pass = true;
      o.dis = (float)sd;
      o.prob = need_prob ? 1.0 / sqrt(sigma_l) * exp(-0.5 * (double)dis_to_plane * (double)dis_to_plane / sigma_l) : 1.0;
    }
  }
  return o;
}

__device__ __forceinline__ bool probe(const HashSlot *__restrict__ slots, uint32_t mask, long long kx, long long ky, long long kz,
                                      uint32_t &first, uint32_t &count) {
  if (!key_in_range(kx, ky, kz)) return false;
  unsigned long long key = pack_key(kx, ky, kz);
  uint32_t s = hash_key(key) & mask;
  for (;;) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(slots + s));
    if (v.x == key) {
      first = (uint32_t)(v.y & 0xffffffffull);
      count = (uint32_t)(v.y >> 32);
      return true;
    }
    if (v.x == ESIKF_KEY_EMPTY) return false;
    s = (s + 1) & mask;
  }
}

// Extra candidates (sub-divided root voxels) of ALL lanes of the warp that have some, in one pass: the (owner lane,
// candidate) pairs are laid out consecutively and dealt one per lane, so the scattered plane-record reads of every pending
// point overlap instead of being paid once per pending lane (the slowest warp of the slowest CTA sets the grid barrier).
// Winner per owner = arg-max probability with lowest-index tie break, merged with `best` by strict '>' — exactly the
// order-dependent rule of the recursion (voxel_map.cpp:741: the first of equal probabilities is kept).
__device__ __forceinline__ void warp_eval_extras(const esikf_plane *__restrict__ planes, bool pending, const double pw[3], const PointCovRef &pc,
                                                 int point, uint32_t first, uint32_t count, double sigma_num, int lane, Cand &best) {
  const unsigned mask = __ballot_sync(0xffffffffu, pending);
  if (!mask) return;
  const int npairs = pending ? (int)count - 1 : 0;
  int scan = npairs;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, scan, d);
    if (lane >= d) scan += t;
  }
  const int total = __shfl_sync(0xffffffffu, scan, 31);
  const int excl = scan - npairs;
  Cand acc;
  acc.prob = -1.0, acc.idx = -1, acc.dis = 0.f;
  for (int base = 0; base < total; base += 32) {
    const int k = base + lane;
    int owner = 0, cand = 0;
    for (unsigned m = mask; m; m &= m - 1) {
      const int jl = __ffs(m) - 1;
      const int ej = __shfl_sync(0xffffffffu, excl, jl), nj = __shfl_sync(0xffffffffu, npairs, jl);
      if (k >= ej && k < ej + nj) owner = jl, cand = k - ej + 1;
    }
    const bool have = k < total;
    double opw[3];
#pragma unroll
    for (int c = 0; c < 3; c++) opw[c] = __shfl_sync(0xffffffffu, pw[c], owner);
    const int opoint = __shfl_sync(0xffffffffu, point, owner);
    const uint32_t ofirst = __shfl_sync(0xffffffffu, first, owner);
    Cand my;
    my.prob = -1.0, my.idx = 0x7fffffff, my.dis = 0.f;
    if (have) {
      PointCovRef opc = pc;
      opc.pre += opoint - point;
      const EvalOut e = eval_rec(reinterpret_cast<const double *>(planes + ofirst + cand), opw, opc, sigma_num, true);
      if (e.pass) my.prob = e.prob, my.idx = (int)(ofirst + cand), my.dis = e.dis;
    }
    for (unsigned m = mask; m; m &= m - 1) {
      const int jl = __ffs(m) - 1;
      const bool mine = have && owner == jl && my.idx != 0x7fffffff;
      double rp = mine ? my.prob : -1.0;
      int ri = mine ? my.idx : 0x7fffffff;
      float rd = mine ? my.dis : 0.f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double op = __shfl_xor_sync(0xffffffffu, rp, off);
        const int oi = __shfl_xor_sync(0xffffffffu, ri, off);
        const float od = __shfl_xor_sync(0xffffffffu, rd, off);
        if (op > rp || (op == rp && oi < ri)) rp = op, ri = oi, rd = od;
      }
      // chunks are visited in increasing candidate order: a later chunk only replaces on strictly larger probability
      if (lane == jl && ri != 0x7fffffff && rp > acc.prob) acc.prob = rp, acc.idx = ri, acc.dis = rd;
    }
  }
  if (pending && acc.idx >= 0 && acc.prob > best.prob) best = acc;
}

#define REC_STRIDE 38  // doubles per lane slot (304 B = 19 x 16 B: conflict-free 128-bit reads at lane stride)
#define REC_ROW 28     // slot layout: [0,28) staged plane record | [28,35) row a_i(7) | 35 R_inv | 36 {f32 |dis|, u32 count} | 37 packed voxel key

// Per-lane state that survives from one iteration to the next inside the persistent kernel (a lane keeps its point).
struct LaneCache {
  int staged_idx;   // plane whose record is resident in the lane's slot (-1: none); its voxel key / candidate count sit in the slot tail
  bool have_pt;
  float px, py, pz; // the body-frame point
};

// shared-memory layout of the residual kernel
struct __align__(128) LioSmem {
  // Per-lane slot: the first candidate plane of the lane's point, staged by coalesced warp copies and kept resident
  // across the iterations of the persistent kernel (re-copied only when the point changes voxel), followed by the
  // lane's row a_i = [A(3) n(3) z 1], R_inv and |dis_to_plane| for the tensor-core contraction.
  double rec[LIO_WARPS][32][REC_STRIDE];
  double R[9], t[3], Ptt[9], Ppp[9];      // current state
  double Rp[9], tp[3], Mp[9];             // prior pose, Mp = Rp * extR
  ReduceSmem<LIO_WARPS> red;
  unsigned char fs_raw[6400];             // CTA 0's resident solve scratch (FusedSolveSmem) in the persistent kernel
};

// Load the per-iteration constants (current pose / covariance blocks, prior pose) into shared memory.
__device__ __forceinline__ void lio_load_consts(LioSmem &sm, const LioKernelArgs &a) {
  const int tid = threadIdx.x;
  if (tid < 9) {
    sm.R[tid] = __ldcg(a.state + S_R + tid);
    sm.Rp[tid] = a.prop[S_R + tid];
    int r = tid / 3, c = tid % 3;
    sm.Ptt[tid] = __ldcg(a.state + S_COV + r * 19 + c);
    sm.Ppp[tid] = __ldcg(a.state + S_COV + (3 + r) * 19 + (3 + c));
    // Mp = Rp * extR  (state_propagat.rot_end * extR_, voxel_map.cpp:445)
    double s = 0;
    for (int k = 0; k < 3; k++) s += a.prop[S_R + r * 3 + k] * a.extR[k * 3 + c];
    sm.Mp[tid] = s;
  } else if (tid < 12) {
    sm.t[tid - 9] = __ldcg(a.state + S_P + tid - 9);
    sm.tp[tid - 9] = a.prop[S_P + tid - 9];
  }
  __syncthreads();
}

// Residual / Jacobian build over the points [lo, hi) of this rank's shard (indices local to the shard), accumulated into
// the calling warp's 8x8 tensor-core block (D0, D1) and matched-point count.
// DEAL = false: the CTA owns the contiguous block [lo, hi) and walks it in tiles of LIO_THREADS points.
// DEAL = true (lo = 0, hi = points of the shard): 32-point chunks are dealt round-robin over the CTAs — chunk
// (t * LIO_WARPS + warp) * gridDim.x + blockIdx.x belongs to this warp in tile t — so a CTA samples the whole scan instead
// of inheriting the local structure (sub-divided voxels, unmatched regions) of one stretch of it. Loads stay coalesced per
// warp; only the fixed summation order differs between the two schedules.
template <bool DEAL = false>
__device__ __forceinline__ void lio_process_range(const LioKernelArgs &a, LioSmem &sm, int lo, int hi, double &D0, double &D1, int &cnt,
                                                  LaneCache &lc, bool init_normal) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *const myrec = &sm.rec[warp][lane][0];
  const int tile_pts = DEAL ? (int)(LIO_THREADS * gridDim.x) : LIO_THREADS;
  for (int base = lo; base < hi; base += tile_pts) {
    const int li = DEAL ? base + (warp * (int)gridDim.x + (int)blockIdx.x) * 32 + lane : base + tid;
    const bool valid = li < hi;
    const int i = a.begin + li;
    if (hi - lo > tile_pts) lc.staged_idx = -1, lc.have_pt = false;  // several tiles share the lanes: nothing stays resident
    int midx = -1;
    float mdis = 0.f;
    double pw[3] = {0, 0, 0};
    float loc[3] = {0, 0, 0};
    uint32_t first = 0, count = 0;
    bool found = false;

    const bool first_tile = (base == lo);
    if (first_tile) dbg_stamp(a.dbg, 0);
    // ---- phase 1: transform, voxel key, home-voxel probe
    if (valid) {
      if (!lc.have_pt) {
        lc.px = a.pts[3 * (size_t)i], lc.py = a.pts[3 * (size_t)i + 1], lc.pz = a.pts[3 * (size_t)i + 2];
        lc.have_pt = (hi - lo <= tile_pts);
      }
      const double px = lc.px, py = lc.py, pz = lc.pz;
      // p_imu = extR p + extT ; p_w = R p_imu + t, narrowed to float (TransformLidar, voxel_map.cpp:522-526). No FMA contraction
      // on this chain: the float rounding of p_w decides the voxel key.
      const double pi0 = __dadd_rn(dot3_rn(a.extR[0], a.extR[1], a.extR[2], px, py, pz), a.extT[0]);
      const double pi1 = __dadd_rn(dot3_rn(a.extR[3], a.extR[4], a.extR[5], px, py, pz), a.extT[1]);
      const double pi2 = __dadd_rn(dot3_rn(a.extR[6], a.extR[7], a.extR[8], px, py, pz), a.extT[2]);
      pw[0] = (double)(float)__dadd_rn(dot3_rn(sm.R[0], sm.R[1], sm.R[2], pi0, pi1, pi2), sm.t[0]);
      pw[1] = (double)(float)__dadd_rn(dot3_rn(sm.R[3], sm.R[4], sm.R[5], pi0, pi1, pi2), sm.t[1]);
      pw[2] = (double)(float)__dadd_rn(dot3_rn(sm.R[6], sm.R[7], sm.R[8], pi0, pi1, pi2), sm.t[2]);
      // voxel key (voxel_map.cpp:665-671): float quotient, "-1 if negative", truncate. When 1/voxel_size is exact (a power
      // of two: 0.5, 2.0, ...) the multiply gives the bit-identical quotient without the slow fp64 division.
      bool finite = true;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        loc[j] = a.inv_voxel_exact ? (float)__dmul_rn(pw[j], a.inv_voxel_size) : (float)__ddiv_rn(pw[j], a.voxel_size);
        if (loc[j] < 0) loc[j] = (float)__dadd_rn((double)loc[j], -1.0);
        finite = finite && (fabsf(loc[j]) < 3.0e6f);
      }
      const long long key[3] = {(long long)loc[0], (long long)loc[1], (long long)loc[2]};
      // the voxel of the previous iteration (key + candidate range cached in the slot tail) needs no second hash probe
      if (finite && lc.staged_idx >= 0 && key_in_range(key[0], key[1], key[2]) &&
          pack_key(key[0], key[1], key[2]) == *reinterpret_cast<const unsigned long long *>(myrec + 37)) {
        found = true;
        first = (uint32_t)lc.staged_idx;
        count = reinterpret_cast<const uint32_t *>(myrec + 36)[1];
      } else {
        found = finite && probe(a.slots, a.hash_mask, key[0], key[1], key[2], first, count);
      }
    }

    if (first_tile) dbg_stamp(a.dbg, 1);
    // ---- phase 2: stage every lane's first candidate record (224 B) with coalesced half-warp copies. A record already
    // resident in the lane's slot (same plane as in the previous iteration of the persistent kernel) is not re-read.
    {
      const int cand0 = (found && count > 0) ? (int)first : -1;
      const int want = (cand0 >= 0 && cand0 != lc.staged_idx) ? cand0 : -1;
      const int half = lane >> 4, sub = lane & 15;
      if (__any_sync(0xffffffffu, want >= 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int src = j + half;
          const int pidx = __shfl_sync(0xffffffffu, want, src);
          if (pidx >= 0 && sub < 14) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(a.planes + pidx) + sub);
            *reinterpret_cast<double2 *>(&sm.rec[warp][src][2 * sub]) = v;
          }
        }
      }
      lc.staged_idx = cand0;
      __syncwarp();
    }

    if (first_tile) dbg_stamp(a.dbg, 2);
    // ---- phase 3: association
    Cand best;
    best.prob = 0.0, best.idx = -1, best.dis = 0.f;
    PointCovRef pc;
    pc.pre = a.pre + i, pc.ns = (size_t)a.pre_stride, pc.R = sm.R, pc.Ptt = sm.Ptt, pc.Ppp = sm.Ppp;
    if (found && count > 0) {
      const EvalOut e = eval_rec(myrec, pw, pc, a.sigma_num, count > 1);
      if (e.pass) best.prob = e.prob, best.idx = (int)first, best.dis = e.dis;
    }
    if (first_tile) dbg_stamp(a.dbg, 3);
    // further candidates of sub-divided root voxels (rare): one point at a time, lane-parallel over its candidate list
    warp_eval_extras(a.planes, found && count > 1, pw, pc, i, first, count, a.sigma_num, lane, best);
    if (first_tile) dbg_stamp(a.dbg, 4);
    // one neighbour voxel when the home voxel gave nothing (voxel_map.cpp:680-691). loc is in voxel units, centre / quarter
    // length in metres: reproduced literally.
    uint32_t f2 = 0, c2 = 0;
    bool found2 = false;
    if (found && best.idx < 0) {
      const double vsf = (double)a.voxel_size_f;
      const double ql = (double)(a.voxel_size_f / 4.0f);
      const long long key[3] = {(long long)loc[0], (long long)loc[1], (long long)loc[2]};
      long long nk[3] = {key[0], key[1], key[2]};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double center = (0.5 + (double)key[j]) * vsf;
        if ((double)loc[j] > center + ql) nk[j] = key[j] + 1;
        else if ((double)loc[j] < center - ql) nk[j] = key[j] - 1;
      }
      found2 = probe(a.slots, a.hash_mask, nk[0], nk[1], nk[2], f2, c2) && c2 > 0;
      if (found2) {
        const EvalOut e = eval_rec(reinterpret_cast<const double *>(a.planes + f2), pw, pc, a.sigma_num, c2 > 1);
        if (e.pass) best.prob = e.prob, best.idx = (int)f2, best.dis = e.dis;
      }
    }
    warp_eval_extras(a.planes, found2 && c2 > 1, pw, pc, i, f2, c2, a.sigma_num, lane, best);

    if (first_tile) dbg_stamp(a.dbg, 5);
    // ---- phase 4: Jacobian / measurement-noise loop (voxel_map.cpp:414-458) for matched points
    double row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double wgt = 0.0, absd = 0.0;
    const bool matched = best.idx >= 0;
    if (matched) {
      midx = best.idx, mdis = best.dis;
      const double *__restrict__ q = (best.idx == (int)first && count > 0 && found) ? myrec : reinterpret_cast<const double *>(a.planes + best.idx);
      const double2 *__restrict__ q2 = reinterpret_cast<const double2 *>(q);
      const double2 a0 = q2[0], a1 = q2[1], a2 = q2[2];
      const double c0 = a0.x, c1 = a0.y, c2_ = a1.x, n0 = a1.y, n1 = a2.x, n2 = a2.y;
      // p_imu recomputed (cheaper than keeping it live across the association) ; point_world with the PRIOR pose (:425)
      const double px = lc.px, py = lc.py, pz = lc.pz;
      const double pi0 = __dadd_rn(dot3_rn(a.extR[0], a.extR[1], a.extR[2], px, py, pz), a.extT[0]);
      const double pi1 = __dadd_rn(dot3_rn(a.extR[3], a.extR[4], a.extR[5], px, py, pz), a.extT[1]);
      const double pi2 = __dadd_rn(dot3_rn(a.extR[6], a.extR[7], a.extR[8], px, py, pz), a.extT[2]);
      const double w0 = sm.Rp[0] * pi0 + sm.Rp[1] * pi1 + sm.Rp[2] * pi2 + sm.tp[0];
      const double w1 = sm.Rp[3] * pi0 + sm.Rp[4] * pi1 + sm.Rp[5] * pi2 + sm.tp[1];
      const double w2 = sm.Rp[6] * pi0 + sm.Rp[7] * pi1 + sm.Rp[8] * pi2 + sm.tp[2];
      const double J[6] = {w0 - c0, w1 - c1, w2 - c2_, -n0, -n1, -n2};
      const double sigma_l = quad6(q + 6, J);
      // n^T (Mp body_cov Mp^T) n = m^T body_cov m, m = Mp^T n   (:445-449)
      const double m0 = sm.Mp[0] * n0 + sm.Mp[3] * n1 + sm.Mp[6] * n2;
      const double m1 = sm.Mp[1] * n0 + sm.Mp[4] * n1 + sm.Mp[7] * n2;
      const double m2 = sm.Mp[2] * n0 + sm.Mp[5] * n1 + sm.Mp[8] * n2;
      const size_t ns = (size_t)a.pre_stride;
      const double *__restrict__ pre = a.pre + i;
      const double bc[6] = {pre[3 * ns], pre[4 * ns], pre[5 * ns], pre[6 * ns], pre[7 * ns], pre[8 * ns]};
      const double nvn = quad3_sym(bc, m0, m1, m2);
      wgt = 1.0 / (0.001 + sigma_l + nvn);
      // A = [p_imu]x R^T n with the CURRENT rotation (:453)
      const double g0 = sm.R[0] * n0 + sm.R[3] * n1 + sm.R[6] * n2;
      const double g1 = sm.R[1] * n0 + sm.R[4] * n1 + sm.R[7] * n2;
      const double g2 = sm.R[2] * n0 + sm.R[5] * n1 + sm.R[8] * n2;
      row[0] = -pi2 * g1 + pi1 * g2;
      row[1] = pi2 * g0 - pi0 * g2;
      row[2] = -pi1 * g0 + pi0 * g1;
      row[3] = n0, row[4] = n1, row[5] = n2;
      row[6] = -(double)best.dis;  // meas_vec (:457)
      row[7] = 1.0;
      absd = fabs((double)best.dis);
    }
    if (valid) {
      a.match_plane[i] = midx;    // ptpl_list_ membership of this iteration
      a.dis_to_plane[i] = mdis;   // PointToPlane::dis_to_plane_ of this iteration (0 when unmatched)
      if (matched) a.normal_plane[i] = midx;  // pv.normal = plane.normal_ (:744), sticky across iterations
      else if (init_normal) a.normal_plane[i] = -1;
    }
    cnt += __popc(__ballot_sync(0xffffffffu, matched));

    if (first_tile) dbg_stamp(a.dbg, 6);
    // ---- phase 5: stage the 32 rows of this warp and contract them on the fp64 tensor path
    __syncwarp();
    {
      double2 *dst = reinterpret_cast<double2 *>(myrec + REC_ROW);
      dst[0] = make_double2(row[0], row[1]);
      dst[1] = make_double2(row[2], row[3]);
      dst[2] = make_double2(row[4], row[5]);
      union {
        double d;
        struct { float f; uint32_t u; } s;
      } pk;
      pk.s.f = (float)absd, pk.s.u = count;
      union {
        double d;
        unsigned long long u;
      } kk;
      {
        const long long k0 = (long long)loc[0], k1 = (long long)loc[1], k2 = (long long)loc[2];
        kk.u = (found && key_in_range(k0, k1, k2)) ? pack_key(k0, k1, k2) : ESIKF_KEY_EMPTY;
      }
      dst[3] = make_double2(row[6], wgt);
      dst[4] = make_double2(pk.d, kk.d);
    }
    __syncwarp();
    {
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const double *r = &sm.rec[warp][4 * s + t][REC_ROW];
        const double wv = r[7];
        const double v = (g == 7) ? ((wv != 0.0) ? 1.0 : 0.0) : r[g];  // a_7 = 1 for matched rows (R_inv > 0), else 0
        const double b = (g == 7) ? (double)reinterpret_cast<const float *>(r + 8)[0] : wv * v;
        dmma_m8n8k4(D0, D1, v, b);
      }
    }
    __syncwarp();
  }

}

// Contiguous, equal slices of the shard per block: every SM gets the same number of points.
__device__ __forceinline__ void lio_block_range(int count, int &lo, int &hi) {
  const int per = (count + gridDim.x - 1) / gridDim.x;
  lo = blockIdx.x * per;
  hi = lo + per < count ? lo + per : count;
  if (lo > count) lo = count;
}

__global__ void __launch_bounds__(LIO_THREADS, 1) lio_residual_kernel(const LioKernelArgs a) {
  if (a.ctrl->stop) return;  // EKF_stop_flg: remaining iterations of the unrolled loop do nothing
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LioSmem &sm = *reinterpret_cast<LioSmem *>(smem_raw);
  lio_load_consts(sm, a);
  double D0 = 0.0, D1 = 0.0;  // this lane's two entries of the warp's 8x8 block
  int cnt = 0;
  int lo, hi;
  lio_block_range(a.count, lo, hi);
  LaneCache lc;
  lc.staged_idx = -1, lc.have_pt = false, lc.px = lc.py = lc.pz = 0.f;
  lio_process_range(a, sm, lo, hi, D0, D1, cnt, lc, a.init_normal != 0);
  reduce_info<LIO_WARPS>(sm.red, D0, D1, (double)cnt, true, a.partials, a.partial_stride, a.info, a.ctrl);
}

}  // namespace esikf
