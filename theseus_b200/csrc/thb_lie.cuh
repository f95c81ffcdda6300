// SO3 / SE3 closed forms as register-resident __device__ functions (one thread = one group element).
//
// Numerical conventions follow the reference's torchlie kernels so that every near-zero / near-pi
// branch is taken on exactly the same inputs:
//   torchlie/torchlie/functional/so3_impl.py  (_exp_impl_helper :220-261, _log_impl_helper :390-433,
//                                              _jlog_impl_helper :442-479)
//   torchlie/torchlie/functional/se3_impl.py  (_exp_impl_helper :178-216, _log_impl_helper :354-396,
//                                              _jlog_impl_helper :405-457, _adjoint_impl :531-538,
//                                              _inverse_impl :578-581, _compose_impl :703-708)
//   eps table: torchlie/torchlie/global_params.py:44-58
// Storage: SO3 row-major 3x3 (9 scalars); SE3 row-major 3x4 = [R|t] (12 scalars);
// SE3 tangent = [v(3), w(3)] (translation first).  No re-normalisation anywhere (the reference
// never projects back onto SO(3) inside the optimisation loop).
#pragma once
#include <cuda_runtime.h>

namespace thb {

template <typename T> struct LieEps;
template <> struct LieEps<float> {
  static constexpr float near_zero = 1e-2f, near_pi = 1e-2f, d_near_zero = 2e-1f;
};
template <> struct LieEps<double> {
  static constexpr double near_zero = 5e-3, near_pi = 1e-7, d_near_zero = 1e-2;
};

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ void t_sincos(T x, T* s, T* c);
template <> __device__ __forceinline__ void t_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void t_sincos<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <typename T> __device__ __forceinline__ T t_atan2(T y, T x);
template <> __device__ __forceinline__ float t_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> __device__ __forceinline__ double t_atan2<double>(double y, double x) { return atan2(y, x); }

// ---------------------------------------------------------------------------------------------
// small fixed-size helpers (row-major)
template <typename T> __device__ __forceinline__ void mat3_mul(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

// SE3 (3x4, row stride 4) helpers -----------------------------------------------------------------
// C = A * B   (se3_impl.py:703-708)
template <typename T> __device__ __forceinline__ void se3_compose(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) C[i * 4 + j] = A[i * 4 + 0] * B[0 * 4 + j] + A[i * 4 + 1] * B[1 * 4 + j] + A[i * 4 + 2] * B[2 * 4 + j];
    C[i * 4 + 3] = A[i * 4 + 0] * B[3] + A[i * 4 + 1] * B[7] + A[i * 4 + 2] * B[11] + A[i * 4 + 3];
  }
}
// C = A^-1 = [R^T | -R^T t]   (se3_impl.py:578-581)
template <typename T> __device__ __forceinline__ void se3_inverse(const T* A, T* C) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) C[i * 4 + j] = A[j * 4 + i];
    C[i * 4 + 3] = -(A[0 * 4 + i] * A[3] + A[1 * 4 + i] * A[7] + A[2 * 4 + i] * A[11]);
  }
}
// C = A^-1 * B computed the way the reference does it (inverse first, then compose).
template <typename T> __device__ __forceinline__ void se3_between(const T* A, const T* B, T* C) {
  T Ai[12];
  se3_inverse(A, Ai);
  se3_compose(Ai, B, C);
}

// ---------------------------------------------------------------------------------------------
// SO3 exp pieces shared with SE3 exp.  w[3] -> R (row stride RS), plus the scalar coefficients.
template <typename T> struct So3ExpCoef {
  T theta, theta2, theta_nz, theta2_nz, sine, cosine, sine_by_theta, omc_by_theta2;
  bool near_zero;
};

template <typename T, int RS> __device__ __forceinline__ So3ExpCoef<T> so3_exp(const T* w, T* R) {
  So3ExpCoef<T> c;
  c.theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  c.theta = t_sqrt(c.theta2);
  // The reference computes theta = ||w|| then theta2 = theta**2 (so3_impl.py:221-222).
  c.theta2 = c.theta * c.theta;
  c.near_zero = c.theta < LieEps<T>::near_zero;
  c.theta_nz = c.near_zero ? T(1) : c.theta;
  c.theta2_nz = c.near_zero ? T(1) : c.theta2;
  T s, co;
  t_sincos(c.theta, &s, &co);
  c.sine = s;
  c.cosine = c.near_zero ? (T(8) / (T(4) + c.theta2) - T(1)) : co;
  c.sine_by_theta = c.near_zero ? (T(0.5) * c.cosine + T(0.5)) : (s / c.theta_nz);
  c.omc_by_theta2 = c.near_zero ? (T(0.5) * c.sine_by_theta) : ((T(1) - c.cosine) / c.theta2_nz);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R[i * RS + j] = c.omc_by_theta2 * w[i] * w[j];
  R[0 * RS + 0] += c.cosine;
  R[1 * RS + 1] += c.cosine;
  R[2 * RS + 2] += c.cosine;
  const T sa0 = c.sine_by_theta * w[0], sa1 = c.sine_by_theta * w[1], sa2 = c.sine_by_theta * w[2];
  R[0 * RS + 1] -= sa2;
  R[1 * RS + 0] += sa2;
  R[0 * RS + 2] += sa1;
  R[2 * RS + 0] -= sa1;
  R[1 * RS + 2] -= sa0;
  R[2 * RS + 1] += sa0;
  return c;
}

// SE3 exp: xi[6] = [v, w] -> G[12]   (se3_impl.py:178-216)
template <typename T> __device__ __forceinline__ void se3_exp(const T* xi, T* G) {
  const T* v = xi;
  const T* w = xi + 3;
  So3ExpCoef<T> c = so3_exp<T, 4>(w, G);
  const T theta3_nz = c.theta_nz * c.theta2_nz;
  const T tms = c.near_zero ? (T(1.0 / 6) - c.theta2 / T(120)) : ((c.theta - c.sine) / theta3_nz);
  const T cx = w[1] * v[2] - w[2] * v[1], cy = w[2] * v[0] - w[0] * v[2], cz = w[0] * v[1] - w[1] * v[0];
  const T wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  G[3] = c.sine_by_theta * v[0] + c.omc_by_theta2 * cx + tms * (w[0] * wv);
  G[7] = c.sine_by_theta * v[1] + c.omc_by_theta2 * cy + tms * (w[1] * wv);
  G[11] = c.sine_by_theta * v[2] + c.omc_by_theta2 * cz + tms * (w[2] * wv);
}

// Right Jacobian of the SO3 exponential (so3_impl.py:270-320, _jexp_impl): J[9] row-major,
//   J = sin(t)/t I + (t - sin t)/t^3 w w^T - (1 - cos t)/t^2 [w]x     (near zero: the w w^T coefficient is 0, as in the reference).
// `tms` is the caller's (theta - sine)/theta^3 coefficient for the rotation block.
template <typename T> __device__ __forceinline__ void so3_jexp_from_coef(const T* w, const So3ExpCoef<T>& c, T tms_rot, T* J) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J[i * 3 + j] = tms_rot * w[i] * w[j];
  J[0] += c.sine_by_theta;
  J[4] += c.sine_by_theta;
  J[8] += c.sine_by_theta;
  const T t0 = c.omc_by_theta2 * w[0], t1 = c.omc_by_theta2 * w[1], t2 = c.omc_by_theta2 * w[2];
  J[1] += t2;
  J[3] -= t2;
  J[2] -= t1;
  J[6] += t1;
  J[5] += t0;
  J[7] -= t0;
}

// SO3 exp + right Jacobian: w[3] -> R[9] (row stride 3), J[9].
template <typename T> __device__ __forceinline__ void so3_exp_jexp(const T* w, T* R, T* J) {
  const So3ExpCoef<T> c = so3_exp<T, 3>(w, R);
  const T theta3_nz = c.theta_nz * c.theta2_nz;
  const T tms = c.near_zero ? T(0) : ((c.theta - c.sine) / theta3_nz);
  so3_jexp_from_coef(w, c, tms, J);
}

// SE3 exp + right Jacobian (se3_impl.py:225-330, _jexp_impl_helper / _jexp_impl): xi[6] = [v, w] -> G[12], J[36] row-major with
//   J = [[Jr, R^T Jt], [0, Jr]],  Jr = the SO3 right Jacobian of w,  Jt = d t / d w assembled from the series coefficients below.
template <typename T> __device__ __forceinline__ void se3_exp_jexp(const T* xi, T* G, T* J) {
  const T* v = xi;
  const T* w = xi + 3;
  se3_exp(xi, G);
  T Rtmp[9];
  const So3ExpCoef<T> c = so3_exp<T, 3>(w, Rtmp);
  const T theta3_nz = c.theta_nz * c.theta2_nz;
  const T tms_t = c.near_zero ? (T(1.0 / 6) - c.theta2 / T(120)) : ((c.theta - c.sine) / theta3_nz);
  const T tms_rot = c.near_zero ? T(0) : tms_t;
  T Jr[9];
  so3_jexp_from_coef(w, c, tms_rot, Jr);
  const T d_omc = c.near_zero ? T(-1.0 / 12) : ((c.sine_by_theta - T(2) * c.omc_by_theta2) / c.theta2_nz);
  const T d_tms = c.near_zero ? T(-1.0 / 60) : ((c.omc_by_theta2 - T(3) * tms_t) / c.theta2_nz);
  const T wv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
  const T wwv[3] = {w[1] * wv[2] - w[2] * wv[1], w[2] * wv[0] - w[0] * wv[2], w[0] * wv[1] - w[1] * wv[0]};
  const T sw[3] = {tms_t * w[0], tms_t * w[1], tms_t * w[2]};
  T Jt[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Jt[i * 3 + j] = (d_omc * wv[i] + d_tms * wwv[i]) * w[j] - v[i] * sw[j];
  const T tv[3] = {-c.omc_by_theta2 * v[0] - tms_t * wv[0], -c.omc_by_theta2 * v[1] - tms_t * wv[1], -c.omc_by_theta2 * v[2] - tms_t * wv[2]};
  // + hat(tv)
  Jt[1] -= tv[2];
  Jt[2] += tv[1];
  Jt[3] += tv[2];
  Jt[5] -= tv[0];
  Jt[6] -= tv[1];
  Jt[7] += tv[0];
  const T sv = sw[0] * v[0] + sw[1] * v[1] + sw[2] * v[2];
  Jt[0] += sv;
  Jt[4] += sv;
  Jt[8] += sv;
#pragma unroll
  for (int q = 0; q < 36; q++) J[q] = T(0);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      J[i * 6 + j] = Jr[i * 3 + j];
      J[(i + 3) * 6 + (j + 3)] = Jr[i * 3 + j];
      // R^T Jt, R = rotation block of G (row stride 4)
      J[i * 6 + (j + 3)] = G[0 * 4 + i] * Jt[0 * 3 + j] + G[1 * 4 + i] * Jt[1 * 3 + j] + G[2 * 4 + i] * Jt[2 * 3 + j];
    }
}

// ---------------------------------------------------------------------------------------------
// SO3 log (so3_impl.py:390-433).  R has row stride RS.  Returns theta, sine, cosine for jlog.
template <typename T> struct So3LogAux {
  T theta, sine, cosine;
};

template <typename T, int RS> __device__ __forceinline__ So3LogAux<T> so3_log(const T* R, T* w) {
  So3LogAux<T> a;
  T sa[3];
  sa[0] = T(0.5) * (R[2 * RS + 1] - R[1 * RS + 2]);
  sa[1] = T(0.5) * (R[0 * RS + 2] - R[2 * RS + 0]);
  sa[2] = T(0.5) * (R[1 * RS + 0] - R[0 * RS + 1]);
  a.cosine = T(0.5) * (R[0] + R[RS + 1] + R[2 * RS + 2] - T(1));
  a.sine = t_sqrt(sa[0] * sa[0] + sa[1] * sa[1] + sa[2] * sa[2]);
  a.theta = t_atan2(a.sine, a.cosine);
  const bool near_zero = a.theta < LieEps<T>::near_zero;
  const bool near_pi = (T(1) + a.cosine) <= LieEps<T>::near_pi;
  const bool nzp = near_zero || near_pi;
  const T sine_nz = nzp ? T(1) : a.sine;
  const T scale = nzp ? (T(1) + a.sine * a.sine / T(6)) : (a.theta / sine_nz);
  if (!near_pi) {
    w[0] = sa[0] * scale;
    w[1] = sa[1] * scale;
    w[2] = sa[2] * scale;
  } else {
    // theta ~ pi: pick the major diagonal entry (so3_impl.py:411-430)
    const T d0 = R[0], d1 = R[RS + 1], d2 = R[2 * RS + 2];
    const int major = ((d1 > d0) && (d1 > d2) ? 1 : 0) + 2 * ((d2 > d0) && (d2 > d1) ? 1 : 0);
    T sel[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      T row = (major == 0) ? R[0 * RS + j] : ((major == 1) ? R[1 * RS + j] : R[2 * RS + j]);
      T col = (major == 0) ? R[j * RS + 0] : ((major == 1) ? R[j * RS + 1] : R[j * RS + 2]);
      sel[j] = T(0.5) * (row + col);
    }
    if (major == 0) sel[0] -= a.cosine;
    else if (major == 1) sel[1] -= a.cosine;
    else sel[2] -= a.cosine;
    const T nrm = t_sqrt(sel[0] * sel[0] + sel[1] * sel[1] + sel[2] * sel[2]);
    const T den = near_zero ? T(1) : nrm;
    const T sm = (major == 0) ? sa[0] : ((major == 1) ? sa[1] : sa[2]);
    const T sgn = (sm > T(0)) ? T(1) : ((sm < T(0)) ? T(-1) : T(1));
    const T f = a.theta * sgn;
    w[0] = (sel[0] / den) * f;
    w[1] = (sel[1] / den) * f;
    w[2] = (sel[2] / den) * f;
  }
  return a;
}

// SO3 jlog (so3_impl.py:442-479): J (3x3, row stride JS) from w, theta, sine, cosine; also b*w.
template <typename T, int JS>
__device__ __forceinline__ void so3_jlog(const T* w, const So3LogAux<T>& x, T* J, T* bw) {
  const bool dnz = x.theta < LieEps<T>::d_near_zero;
  const T theta2 = x.theta * x.theta;
  const T st = x.sine * x.theta;
  const T tcm2 = T(2) * x.cosine - T(2);
  const T tcm2_nz = dnz ? T(1) : tcm2;
  const T theta2_nz = dnz ? T(1) : theta2;
  const T a = dnz ? (T(1) - theta2 / T(12)) : (-st / tcm2_nz);
  const T b = dnz ? (T(1.0 / 12) + theta2 / T(720)) : ((st + tcm2) / (theta2_nz * tcm2_nz));
  bw[0] = b * w[0];
  bw[1] = b * w[1];
  bw[2] = b * w[2];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J[i * JS + j] = bw[i] * w[j];
  const T h0 = T(0.5) * w[0], h1 = T(0.5) * w[1], h2 = T(0.5) * w[2];
  J[0 * JS + 1] -= h2;
  J[1 * JS + 0] += h2;
  J[0 * JS + 2] += h1;
  J[2 * JS + 0] -= h1;
  J[1 * JS + 2] -= h0;
  J[2 * JS + 1] += h0;
  J[0 * JS + 0] += a;
  J[1 * JS + 1] += a;
  J[2 * JS + 2] += a;
}

// ---------------------------------------------------------------------------------------------
// SE3 log + jlog (se3_impl.py:354-457).  G[12] -> xi[6]; if J != nullptr also the 6x6 Jacobian (row-major).
template <typename T, bool WITH_J> __device__ __forceinline__ void se3_log_jlog(const T* G, T* xi, T* J) {
  T* lin = xi;
  T* ang = xi + 3;
  So3LogAux<T> x = so3_log<T, 4>(G, ang);
  const bool near_zero = x.theta < LieEps<T>::near_zero;
  const T theta2 = x.theta * x.theta;
  const T st = x.sine * x.theta;
  const T tcm2 = T(2) * x.cosine - T(2);
  const T tcm2_nz = near_zero ? T(1) : tcm2;
  const T theta2_nz = near_zero ? T(1) : theta2;
  const T a = near_zero ? (T(1) - theta2 / T(12)) : (-st / tcm2_nz);
  const T b = near_zero ? (T(1.0 / 12) + theta2 / T(720)) : ((st + tcm2) / (theta2_nz * tcm2_nz));
  const T t0 = G[3], t1 = G[7], t2 = G[11];
  const T cx = ang[1] * t2 - ang[2] * t1, cy = ang[2] * t0 - ang[0] * t2, cz = ang[0] * t1 - ang[1] * t0;
  const T wt = ang[0] * t0 + ang[1] * t1 + ang[2] * t2;
  lin[0] = a * t0 - T(0.5) * cx + b * (ang[0] * wt);
  lin[1] = a * t1 - T(0.5) * cy + b * (ang[1] * wt);
  lin[2] = a * t2 - T(0.5) * cz + b * (ang[2] * wt);
  if (WITH_J) {
    const bool dnz = x.theta < LieEps<T>::d_near_zero;
    T bw[3];
    so3_jlog<T, 6>(ang, x, J, bw);  // top-left 3x3
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        J[(i + 3) * 6 + (j + 3)] = J[i * 6 + j];
        J[(i + 3) * 6 + j] = T(0);
      }
    const T theta_nz = dnz ? T(1) : x.theta;
    const T theta4_nz = theta2_nz * theta2_nz;  // theta2_nz uses the *near_zero* mask (se3_impl.py:366-368,425)
    const T c = dnz ? (T(-1 / 360.0) - theta2 / T(7560.0))
                    : (-(T(2) * tcm2_nz + x.theta * x.sine + theta2) / (theta4_nz * tcm2_nz));
    const T d = dnz ? (T(-1 / 6.0) - theta2 / T(180.0)) : ((x.theta - x.sine) / (theta_nz * tcm2_nz));
    const T e = ang[0] * lin[0] + ang[1] * lin[1] + ang[2] * lin[2];
    const T ce = c * e;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        J[i * 6 + 3 + j] = (ce * ang[i]) * ang[j] + (bw[i] * lin[j] + lin[i] * bw[j]);
    const T ed = e * d;
    J[0 * 6 + 3] += ed;
    J[1 * 6 + 4] += ed;
    J[2 * 6 + 5] += ed;
    const T h0 = T(0.5) * lin[0], h1 = T(0.5) * lin[1], h2 = T(0.5) * lin[2];
    J[0 * 6 + 4] -= h2;
    J[1 * 6 + 3] += h2;
    J[0 * 6 + 5] += h1;
    J[2 * 6 + 3] -= h1;
    J[1 * 6 + 5] -= h0;
    J[2 * 6 + 4] += h0;
  }
}

// Adjoint of an SE3 element, 6x6 row-major: [[R, hat(t) R], [0, R]]  (se3_impl.py:531-538)
template <typename T> __device__ __forceinline__ void se3_adjoint(const T* G, T* Ad) {
  const T t0 = G[3], t1 = G[7], t2 = G[11];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      Ad[i * 6 + j] = G[i * 4 + j];
      Ad[(i + 3) * 6 + 3 + j] = G[i * 4 + j];
      Ad[(i + 3) * 6 + j] = T(0);
    }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    // hat(t) @ R, row by row; zero entries of hat() are kept as explicit products like the reference matmul
    Ad[0 * 6 + 3 + j] = T(0) * G[0 * 4 + j] + (-t2) * G[1 * 4 + j] + t1 * G[2 * 4 + j];
    Ad[1 * 6 + 3 + j] = t2 * G[0 * 4 + j] + T(0) * G[1 * 4 + j] + (-t0) * G[2 * 4 + j];
    Ad[2 * 6 + 3 + j] = (-t1) * G[0 * 4 + j] + t0 * G[1 * 4 + j] + T(0) * G[2 * 4 + j];
  }
}

// SO3 (3x3 row-major, stride 3) group ops ---------------------------------------------------------
template <typename T> __device__ __forceinline__ void so3_between(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}


// ---------------------------------------------------------------------------------------------
// SE2 (theseus/geometry/se2.py): storage [x, y, cos, sin], tangent [ux, uy, theta]; eps theseus/global_params.py:46-59
template <typename T> struct Se2Eps;
template <> struct Se2Eps<float> { static constexpr float near_zero = 3e-2f, d_near_zero = 1e-1f; };
template <> struct Se2Eps<double> { static constexpr double near_zero = 1e-6, d_near_zero = 1e-3; };

// C = A * B (se2.py:318-332)
template <typename T> __device__ __forceinline__ void se2_compose(const T* A, const T* B, T* C) {
  C[0] = A[2] * B[0] - A[3] * B[1] + A[0];
  C[1] = A[3] * B[0] + A[2] * B[1] + A[1];
  C[2] = A[2] * B[2] - A[3] * B[3];
  C[3] = A[3] * B[2] + A[2] * B[3];
}
// C = A^-1 (se2.py:334-339)
template <typename T> __device__ __forceinline__ void se2_inverse(const T* A, T* C) {
  C[0] = -(A[2] * A[0] + A[3] * A[1]);
  C[1] = -(-A[3] * A[0] + A[2] * A[1]);
  C[2] = A[2];
  C[3] = -A[3];
}
template <typename T> __device__ __forceinline__ void se2_between(const T* A, const T* B, T* C) {
  T Ai[4];
  se2_inverse(A, Ai);
  se2_compose(Ai, B, C);
}
// exp (se2.py:239-268)
template <typename T> __device__ __forceinline__ void se2_exp(const T* xi, T* G) {
  const T theta = xi[2];
  T s, c;
  t_sincos(theta, &s, &c);
  const bool small = (theta < T(0) ? -theta : theta) < Se2Eps<T>::near_zero;
  const T theta_nz = small ? T(1) : theta;
  const T sbt = small ? (T(1) - theta * theta / T(6)) : (s / theta_nz);
  const T cmo = small ? (-theta / T(2) + theta * theta * theta / T(24)) : ((c - T(1)) / theta_nz);
  G[0] = sbt * xi[0] + cmo * xi[1];
  G[1] = sbt * xi[1] - cmo * xi[0];
  G[2] = c;
  G[3] = s;
}
// log + jlog (se2.py:165-228); J row-major 3x3
template <typename T, bool WITH_J> __device__ __forceinline__ void se2_log_jlog(const T* G, T* xi, T* J) {
  const T cosine = G[2], sine = G[3];
  const T theta = t_atan2(sine, cosine);
  const T at = theta < T(0) ? -theta : theta;
  const bool small = at < Se2Eps<T>::near_zero;
  const T sine_nz = small ? T(1) : sine;
  const T a = T(0.5) * (T(1) + cosine) * (small ? (T(1) + sine * sine / T(6)) : (theta / sine_nz));
  const T half = T(0.5) * theta;
  const T ux = a * G[0] + half * G[1];
  const T uy = a * G[1] - half * G[0];
  xi[0] = ux;
  xi[1] = uy;
  xi[2] = theta;
  if (WITH_J) {
    const bool dsmall = at < Se2Eps<T>::d_near_zero;
    const T theta_nz = dsmall ? T(1) : theta;
    const T omc_nz = dsmall ? T(1) : (T(1) - cosine);
    const T d = dsmall ? (T(1) - theta * theta / T(12)) : (half * sine / omc_nz);
    const T coeff = dsmall ? (theta / T(12) + theta * theta * theta / T(720)) : (T(1) / theta_nz - T(0.5) * sine / omc_nz);
    J[0] = d;      J[1] = -half;  J[2] = coeff * ux + T(0.5) * uy;
    J[3] = half;   J[4] = d;      J[5] = coeff * uy - T(0.5) * ux;
    J[6] = T(0);   J[7] = T(0);   J[8] = T(1);
  }
}
// adjoint (se2.py:309-316)
template <typename T> __device__ __forceinline__ void se2_adjoint(const T* G, T* Ad) {
  Ad[0] = G[2]; Ad[1] = -G[3]; Ad[2] = G[1];
  Ad[3] = G[3]; Ad[4] = G[2];  Ad[5] = -G[0];
  Ad[6] = T(0); Ad[7] = T(0);  Ad[8] = T(1);
}

}  // namespace thb
