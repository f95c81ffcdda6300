// Stand-alone SO3 / SE2 group operations (one thread per element): exp, log (+ jlog), adjoint, inverse, compose.
// Same device functions as the fused cost / retract kernels (thb_lie.cuh); these entry points mirror the reference's
// torchlie.functional SO3 namespace (torchlie/functional/so3_impl.py:220-261 exp, :390-433 log, :442-479 jlog, :669-672 compose,
// :561-563 inverse, adjoint = R) and theseus.geometry.SE2 (theseus/geometry/se2.py:239-300 exp_map, :165-228 log_map + Jacobian,
// :309-316 adjoint, :318-332 compose, :334-339 inverse).  SE3: thb_costs.cu.
#include "thb_common.cuh"
#include "thb_lie.cuh"

namespace thb {

template <typename T> __global__ void k_so3_exp(const T* __restrict__ w, T* __restrict__ R, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T x[3], r[9];
#pragma unroll
  for (int q = 0; q < 3; q++) x[q] = w[i * 3 + q];
  so3_exp<T, 3>(x, r);
#pragma unroll
  for (int q = 0; q < 9; q++) R[i * 9 + q] = r[q];
}
template <typename T> __global__ void k_so3_log(const T* __restrict__ R, T* __restrict__ w, T* __restrict__ J, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T r[9], x[3];
#pragma unroll
  for (int q = 0; q < 9; q++) r[q] = R[i * 9 + q];
  const So3LogAux<T> a = so3_log<T, 3>(r, x);
#pragma unroll
  for (int q = 0; q < 3; q++) w[i * 3 + q] = x[q];
  if (J != nullptr) {
    T jl[9], bw[3];
    so3_jlog<T, 3>(x, a, jl, bw);
#pragma unroll
    for (int q = 0; q < 9; q++) J[i * 9 + q] = jl[q];
  }
}
// mode 0: adjoint (= R), 1: inverse (= R^T)
template <typename T> __global__ void k_so3_unary(const T* __restrict__ R, T* __restrict__ O, int64_t N, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) O[i * 9 + r * 3 + c] = mode == 0 ? R[i * 9 + r * 3 + c] : R[i * 9 + c * 3 + r];
}
template <typename T> __global__ void k_so3_compose(const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ O, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T a[9], b[9];
#pragma unroll
  for (int q = 0; q < 9; q++) { a[q] = A[i * 9 + q]; b[q] = Bm[i * 9 + q]; }
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) O[i * 9 + r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}

template <typename T> __global__ void k_se2_exp(const T* __restrict__ xi, T* __restrict__ G, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T x[3], g[4];
#pragma unroll
  for (int q = 0; q < 3; q++) x[q] = xi[i * 3 + q];
  se2_exp(x, g);
#pragma unroll
  for (int q = 0; q < 4; q++) G[i * 4 + q] = g[q];
}
template <typename T> __global__ void k_se2_log(const T* __restrict__ G, T* __restrict__ xi, T* __restrict__ J, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T g[4], x[3], jl[9];
#pragma unroll
  for (int q = 0; q < 4; q++) g[q] = G[i * 4 + q];
  if (J != nullptr) {
    se2_log_jlog<T, true>(g, x, jl);
#pragma unroll
    for (int q = 0; q < 9; q++) J[i * 9 + q] = jl[q];
  } else {
    se2_log_jlog<T, false>(g, x, jl);
  }
#pragma unroll
  for (int q = 0; q < 3; q++) xi[i * 3 + q] = x[q];
}
template <typename T> __global__ void k_se2_adjoint(const T* __restrict__ G, T* __restrict__ A, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T g[4], ad[9];
#pragma unroll
  for (int q = 0; q < 4; q++) g[q] = G[i * 4 + q];
  se2_adjoint(g, ad);
#pragma unroll
  for (int q = 0; q < 9; q++) A[i * 9 + q] = ad[q];
}
template <typename T> __global__ void k_se2_inverse(const T* __restrict__ G, T* __restrict__ O, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T g[4], o[4];
#pragma unroll
  for (int q = 0; q < 4; q++) g[q] = G[i * 4 + q];
  se2_inverse(g, o);
#pragma unroll
  for (int q = 0; q < 4; q++) O[i * 4 + q] = o[q];
}
template <typename T> __global__ void k_se2_compose(const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ O, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T a[4], b[4], o[4];
#pragma unroll
  for (int q = 0; q < 4; q++) { a[q] = A[i * 4 + q]; b[q] = Bm[i * 4 + q]; }
  se2_compose(a, b, o);
#pragma unroll
  for (int q = 0; q < 4; q++) O[i * 4 + q] = o[q];
}

// exp + right Jacobian of exp (so3_impl.py:270-320, se3_impl.py:225-330); `group` may be NULL
template <typename T> __global__ void k_so3_jexp(const T* __restrict__ w, T* __restrict__ R, T* __restrict__ J, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T x[3], r[9], j[9];
#pragma unroll
  for (int q = 0; q < 3; q++) x[q] = w[i * 3 + q];
  so3_exp_jexp(x, r, j);
#pragma unroll
  for (int q = 0; q < 9; q++) J[i * 9 + q] = j[q];
  if (R != nullptr) {
#pragma unroll
    for (int q = 0; q < 9; q++) R[i * 9 + q] = r[q];
  }
}
template <typename T> __global__ void k_se3_jexp(const T* __restrict__ xi, T* __restrict__ G, T* __restrict__ J, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T x[6], g[12], j[36];
#pragma unroll
  for (int q = 0; q < 6; q++) x[q] = xi[i * 6 + q];
  se3_exp_jexp(x, g, j);
#pragma unroll
  for (int q = 0; q < 36; q++) J[i * 36 + q] = j[q];
  if (G != nullptr) {
#pragma unroll
    for (int q = 0; q < 12; q++) G[i * 12 + q] = g[q];
  }
}

}  // namespace thb

#define THB_OPS_ENTRY(KERNEL, ...)                                                 \
  {                                                                                \
    if (N <= 0) return THB_OK;                                                     \
    KERNEL<<<(unsigned)((N + 127) / 128), 128, 0, thb_cs(s)>>>(__VA_ARGS__);       \
    THB_CHECK_LAUNCH();                                                            \
    return THB_OK;                                                                 \
  }

extern "C" {
#define THB_OPS_FOR(T, SFX)                                                                                                                   \
  int thb_so3_exp_##SFX(const T* t, T* g, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_exp<T>, t, g, N)                                \
  int thb_so3_log_##SFX(const T* g, T* t, T* j, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_log<T>, g, t, j, N)                       \
  int thb_so3_adjoint_##SFX(const T* g, T* a, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_unary<T>, g, a, N, 0)                       \
  int thb_so3_inverse_##SFX(const T* g, T* o, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_unary<T>, g, o, N, 1)                       \
  int thb_so3_compose_##SFX(const T* a, const T* b, T* o, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_compose<T>, a, b, o, N)         \
  int thb_se2_exp_##SFX(const T* t, T* g, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se2_exp<T>, t, g, N)                                \
  int thb_se2_log_##SFX(const T* g, T* t, T* j, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se2_log<T>, g, t, j, N)                       \
  int thb_se2_adjoint_##SFX(const T* g, T* a, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se2_adjoint<T>, g, a, N)                        \
  int thb_se2_inverse_##SFX(const T* g, T* o, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se2_inverse<T>, g, o, N)                        \
  int thb_se2_compose_##SFX(const T* a, const T* b, T* o, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se2_compose<T>, a, b, o, N)         \
  int thb_so3_jexp_##SFX(const T* t, T* g, T* j, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_so3_jexp<T>, t, g, j, N)                     \
  int thb_se3_jexp_##SFX(const T* t, T* g, T* j, int64_t N, thb_stream_t s) THB_OPS_ENTRY(thb::k_se3_jexp<T>, t, g, j, N)
THB_OPS_FOR(double, f64)
THB_OPS_FOR(float, f32)
}  // extern "C"
