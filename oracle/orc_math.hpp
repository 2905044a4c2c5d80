// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path
// (fast_livo2_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// PARITY UNPINNED: the reference (hku-mars/FAST-LIVO2) ships no tests, golden vectors or
// fixtures for this path and cannot be compiled in this image (needs ROS1/PCL/Eigen/Sophus/
// OpenCV/vikit). This is a dependency-free restatement of the reference's algorithm.
//
// Small fixed-size dense helpers standing in for Eigen (absent from the image). Row-major.
// Evaluation order of products is plain left-to-right dot products; build the parity
// library with -ffp-contract=off so no FMA contraction happens.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

template <int R, int C> struct Mat {
  double a[R * C];
  double &operator()(int r, int c) { return a[r * C + c]; }
  const double &operator()(int r, int c) const { return a[r * C + c]; }
  double &operator[](int i) { return a[i]; }
  const double &operator[](int i) const { return a[i]; }
  static Mat Zero() {
    Mat m;
    for (int i = 0; i < R * C; i++) m.a[i] = 0.0;
    return m;
  }
  static Mat Identity() {
    Mat m = Zero();
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
};

typedef Mat<3, 3> M3;
typedef Mat<3, 1> V3;
typedef Mat<2, 1> V2;
typedef Mat<2, 2> M2;
typedef Mat<19, 19> M19;
typedef Mat<19, 1> V19;

template <int R, int K, int C> inline Mat<R, C> operator*(const Mat<R, K> &A, const Mat<K, C> &B) {
  Mat<R, C> o;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      double s = A(r, 0) * B(0, c);
      for (int k = 1; k < K; k++) s = s + A(r, k) * B(k, c);
      o(r, c) = s;
    }
  return o;
}
template <int R, int C> inline Mat<R, C> operator+(const Mat<R, C> &A, const Mat<R, C> &B) {
  Mat<R, C> o;
  for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] + B.a[i];
  return o;
}
template <int R, int C> inline Mat<R, C> operator-(const Mat<R, C> &A, const Mat<R, C> &B) {
  Mat<R, C> o;
  for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] - B.a[i];
  return o;
}
template <int R, int C> inline Mat<R, C> operator-(const Mat<R, C> &A) {
  Mat<R, C> o;
  for (int i = 0; i < R * C; i++) o.a[i] = -A.a[i];
  return o;
}
template <int R, int C> inline Mat<R, C> operator*(const Mat<R, C> &A, double s) {
  Mat<R, C> o;
  for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] * s;
  return o;
}
template <int R, int C> inline Mat<R, C> operator*(double s, const Mat<R, C> &A) { return A * s; }
template <int R, int C> inline Mat<R, C> operator/(const Mat<R, C> &A, double s) {
  Mat<R, C> o;
  for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] / s;
  return o;
}
template <int R, int C> inline Mat<C, R> T(const Mat<R, C> &A) {
  Mat<C, R> o;
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) o(c, r) = A(r, c);
  return o;
}
template <int N> inline double dot(const Mat<N, 1> &a, const Mat<N, 1> &b) {
  double s = a[0] * b[0];
  for (int i = 1; i < N; i++) s = s + a[i] * b[i];
  return s;
}
template <int N> inline double norm(const Mat<N, 1> &a) { return std::sqrt(dot(a, a)); }
inline V3 v3(double x, double y, double z) {
  V3 v;
  v[0] = x, v[1] = y, v[2] = z;
  return v;
}
inline V3 cross(const V3 &a, const V3 &b) {
  return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// SKEW_SYM_MATRX, include/utils/so3_math.h:7
inline M3 skew(const V3 &v) {
  M3 m;
  m(0, 0) = 0.0, m(0, 1) = -v[2], m(0, 2) = v[1];
  m(1, 0) = v[2], m(1, 1) = 0.0, m(1, 2) = -v[0];
  m(2, 0) = -v[1], m(2, 1) = v[0], m(2, 2) = 0.0;
  return m;
}
template <int R, int C, int R2, int C2> inline void set_block(Mat<R, C> &dst, int r0, int c0, const Mat<R2, C2> &src) {
  for (int r = 0; r < R2; r++)
    for (int c = 0; c < C2; c++) dst(r0 + r, c0 + c) = src(r, c);
}
template <int R2, int C2, int R, int C> inline Mat<R2, C2> block(const Mat<R, C> &src, int r0, int c0) {
  Mat<R2, C2> o;
  for (int r = 0; r < R2; r++)
    for (int c = 0; c < C2; c++) o(r, c) = src(r0 + r, c0 + c);
  return o;
}

// Exp(v1,v2,v3): include/utils/so3_math.h:44-58 (identity when |v| <= 1e-5).
inline M3 Exp(double a1, double a2, double a3) {
  double nrm = std::sqrt(a1 * a1 + a2 * a2 + a3 * a3);
  M3 Eye3 = M3::Identity();
  if (nrm > 0.00001) {
    V3 r = v3(a1 / nrm, a2 / nrm, a3 / nrm);
    M3 K = skew(r);
    return Eye3 + std::sin(nrm) * K + (1.0 - std::cos(nrm)) * (K * K);
  }
  return Eye3;
}
// Log(R): include/utils/so3_math.h:61-66.
inline V3 Log(const M3 &R) {
  double tr = R(0, 0) + R(1, 1) + R(2, 2);
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
  V3 K = v3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
  return (std::fabs(theta) < 0.001) ? (0.5 * K) : (0.5 * theta / std::sin(theta) * K);
}

// Stand-in for Eigen's fixed-size .inverse() (PartialPivLU for N > 4): LU with partial
// (row) pivoting, then solve against the identity. Used for the two 19x19 inversions at
// src/voxel_map.cpp:468 and src/vio.cpp:1661.
template <int N> inline Mat<N, N> inverse_pplu(const Mat<N, N> &Ain) {
  Mat<N, N> LU = Ain;
  int perm[N];
  for (int i = 0; i < N; i++) perm[i] = i;
  for (int k = 0; k < N; k++) {
    int piv = k;
    double best = std::fabs(LU(k, k));
    for (int r = k + 1; r < N; r++) {
      double v = std::fabs(LU(r, k));
      if (v > best) best = v, piv = r;
    }
    if (piv != k) {
      for (int c = 0; c < N; c++) {
        double t = LU(k, c);
        LU(k, c) = LU(piv, c);
        LU(piv, c) = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    double d = LU(k, k);
    for (int r = k + 1; r < N; r++) {
      LU(r, k) = LU(r, k) / d;
      double l = LU(r, k);
      for (int c = k + 1; c < N; c++) LU(r, c) = LU(r, c) - l * LU(k, c);
    }
  }
  Mat<N, N> inv;
  for (int col = 0; col < N; col++) {
    double y[N];
    for (int r = 0; r < N; r++) {
      double s = (perm[r] == col) ? 1.0 : 0.0;
      for (int c = 0; c < r; c++) s = s - LU(r, c) * y[c];
      y[r] = s;
    }
    for (int r = N - 1; r >= 0; r--) {
      double s = y[r];
      for (int c = r + 1; c < N; c++) s = s - LU(r, c) * inv(c, col);
      inv(r, col) = s / LU(r, r);
    }
  }
  return inv;
}
// Eigen's 2x2 inverse is the closed form (cofactors / determinant).
inline M2 inverse2(const M2 &A) {
  double det = A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0);
  double id = 1.0 / det;
  M2 o;
  o(0, 0) = A(1, 1) * id, o(0, 1) = -A(0, 1) * id;
  o(1, 0) = -A(1, 0) * id, o(1, 1) = A(0, 0) * id;
  return o;
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi). The reference uses the general
// Eigen::EigenSolver (src/voxel_map.cpp:70); eigenvector sign/order freedom cancels in
// everything the hot path consumes.
inline void eig_sym3(const M3 &A, double evals[3], M3 &evecs) {
  M3 a = A;
  evecs = M3::Identity();
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = std::fabs(a(0, 1)) + std::fabs(a(0, 2)) + std::fabs(a(1, 2));
    double diag = std::fabs(a(0, 0)) + std::fabs(a(1, 1)) + std::fabs(a(2, 2));
    if (off <= 1e-18 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a(p, q) == 0.0) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = evecs(k, p), vkq = evecs(k, q);
          evecs(k, p) = c * vkp - s * vkq;
          evecs(k, q) = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; i++) evals[i] = a(i, i);
}

// StatesGroup: include/common_lib.h:126-223. Error-state order
// [dtheta(0:3) dp(3:6) dinv_expo(6) dv(7:10) dbg(10:13) dba(13:16) dg(16:19)].
struct StatesGroup {
  M3 rot_end;
  V3 pos_end;
  V3 vel_end;
  double inv_expo_time;
  V3 bias_g;
  V3 bias_a;
  V3 gravity;
  M19 cov;
  StatesGroup() {
    rot_end = M3::Identity();
    pos_end = vel_end = bias_g = bias_a = gravity = V3::Zero();
    inv_expo_time = 1.0;
    cov = M19::Identity() * 0.01;  // INIT_COV, common_lib.h:31,137
    cov(6, 6) = 0.00001;           // :138
    for (int i = 10; i < 19; i++) cov(i, i) = 0.00001;  // :139
  }
  // operator+=, common_lib.h:182-192
  void boxplus(const V19 &d) {
    rot_end = rot_end * Exp(d[0], d[1], d[2]);
    for (int i = 0; i < 3; i++) {
      pos_end[i] += d[3 + i];
      vel_end[i] += d[7 + i];
      bias_g[i] += d[10 + i];
      bias_a[i] += d[13 + i];
      gravity[i] += d[16 + i];
    }
    inv_expo_time += d[6];
  }
  // operator-, common_lib.h:194-206 : this (-) b
  V19 boxminus(const StatesGroup &b) const {
    V19 o;
    M3 rotd = T(b.rot_end) * rot_end;
    V3 l = Log(rotd);
    for (int i = 0; i < 3; i++) {
      o[i] = l[i];
      o[3 + i] = pos_end[i] - b.pos_end[i];
      o[7 + i] = vel_end[i] - b.vel_end[i];
      o[10 + i] = bias_g[i] - b.bias_g[i];
      o[13 + i] = bias_a[i] - b.bias_a[i];
      o[16 + i] = gravity[i] - b.gravity[i];
    }
    o[6] = inv_expo_time - b.inv_expo_time;
    return o;
  }
};

// Packed POD crossing the C boundary (fp64, row-major):
// R[9] p[3] inv_expo v[3] bg[3] ba[3] g[3] cov[361]  = 386 doubles.
enum { STATE_PACK = 386 };
inline void unpack_state(const double *s, StatesGroup &x) {
  for (int i = 0; i < 9; i++) x.rot_end.a[i] = s[i];
  for (int i = 0; i < 3; i++) {
    x.pos_end[i] = s[9 + i];
    x.vel_end[i] = s[13 + i];
    x.bias_g[i] = s[16 + i];
    x.bias_a[i] = s[19 + i];
    x.gravity[i] = s[22 + i];
  }
  x.inv_expo_time = s[12];
  for (int i = 0; i < 361; i++) x.cov.a[i] = s[25 + i];
}
inline void pack_state(const StatesGroup &x, double *s) {
  for (int i = 0; i < 9; i++) s[i] = x.rot_end.a[i];
  for (int i = 0; i < 3; i++) {
    s[9 + i] = x.pos_end[i];
    s[13 + i] = x.vel_end[i];
    s[16 + i] = x.bias_g[i];
    s[19 + i] = x.bias_a[i];
    s[22 + i] = x.gravity[i];
  }
  s[12] = x.inv_expo_time;
  for (int i = 0; i < 361; i++) s[25 + i] = x.cov.a[i];
}

}  // namespace orc
