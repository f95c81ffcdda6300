// Batched dense Cholesky factor + solve with fused LM damping, fp64, for sm_100a.
//
// Replaces torch.linalg.cholesky + torch.cholesky_solve + DenseSolver._apply_damping
// (theseus/optimizer/linear/dense_solver.py:38-64,159-161).
//
// Algorithm: left-looking blocked Cholesky over 128-wide block columns.  For block column j one
// launch runs a CTA per (matrix b, row tile i >= j):
//   A. C = sum_{k<j} L[i,k] L[j,k]^T        DMMA (mma.sync m8n8k4 f64) main loop, cp.async 4-stage pipeline,
//                                            operands staged in shared memory with a conflict-free padded stride
//   B. C = (AtA tile, damping fused on the diagonal) - C      (AtA is read exactly once, never modified)
//   C. diagonal CTA: potrf of the 128x128 tile in shared memory, then in-place triangular inverse
//      W = L_jj^-1; stores L_jj and W; releases a per-(b,j) flag
//   D. off-diagonal CTAs: acquire the flag, L[i,j] = C W^T as a second DMMA product (triangular k-range
//      skipped), so the TRSM also runs on the FP64 tensor pipe
// The diagonal CTAs have the lowest block indices of their launch, so they are resident before (or together
// with) the CTAs that wait on them.  The left-looking order keeps every C tile in registers for its whole
// k-loop: L is written once and AtA read once (algorithmic bytes), instead of the n/NB read-modify-write
// sweeps of a right-looking update.
// Solve: x = M^-1 rhs with the stored W_j (triangular solves become mat-vecs), one CTA per matrix, HBM-bound.
//
// FP64 has no tcgen05 kind; the FP64 tensor pipe on sm_100a is reached with mma.sync DMMA
// (measured here: 37.1 TFLOP/s vs 34.2 for DFMA, profiles/r01_ubench_fp64.txt).
#include "thb_common.cuh"

namespace thb {

constexpr int NB = 128;        // tile edge
constexpr int KB = 16;         // k-step of the pipelined product
constexpr int SA = 20;         // smem row stride (doubles) of a [128 x KB] operand tile: 2*SA mod 32 == 8 -> conflict-free DMMA fragment loads
constexpr int SC = 132;        // smem row stride (doubles) of the 128x128 C tile (2*SC mod 32 == 8)
constexpr int STAGES = 4;
constexpr int WSTAGES = 3;
constexpr int CHOL_THREADS = 256;
constexpr int OPER_TILE = NB * SA;                                    // doubles per operand tile
constexpr size_t SMEM_PHASE_A = (size_t)STAGES * 2 * OPER_TILE * 8;   // 163840
constexpr size_t SMEM_PHASE_D = (size_t)(NB * SC + WSTAGES * OPER_TILE) * 8;  // 135168 + 61440
constexpr size_t CHOL_SMEM = SMEM_PHASE_A > SMEM_PHASE_D ? SMEM_PHASE_A : SMEM_PHASE_D;

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// stage one [128 x KB] operand tile: rows row0.. of a row-major matrix with leading dimension ld, columns k0..k0+KB
__device__ __forceinline__ void load_oper_tile(double* dst, const double* __restrict__ src, int64_t ld, int k0, int tid) {
#pragma unroll
  for (int q = 0; q < (NB * KB / 2) / CHOL_THREADS; q++) {
    const int chunk = tid + q * CHOL_THREADS;
    const int row = chunk >> 3, cc = chunk & 7;
    cp_async16(dst + row * SA + cc * 2, src + (int64_t)row * ld + k0 + cc * 2);
  }
}

struct CholArgs {
  const double* AtA;   // [B,n,n]
  const double* alpha; // [B] or null
  const double* beta;  // [B] or null
  double* L;           // [B,np,np]
  double* W;           // [B,nblk,128,128]
  int* flags;          // [B,nblk]
  int32_t* info;       // [B]
  int64_t B, n, np;
  int nblk, j;
};

__global__ void __launch_bounds__(CHOL_THREADS, 1) chol_col_kernel(CholArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x % p.B;
  const int i = p.j + (int)(blockIdx.x / p.B);
  const int j = p.j;
  const bool is_diag = (i == j);
  const int64_t np = p.np;
  double* Lb = p.L + b * np * np;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps -> 64 x 32 warp tiles
  const int lr = lane >> 2, lc = lane & 3;

  // ---------------- phase A: acc = sum_k L[i,k] L[j,k]^T ----------------
  double acc[8][4][2];
#pragma unroll
  for (int mi = 0; mi < 8; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  const int nk = j * (NB / KB);
  const double* Arow = Lb + (int64_t)i * NB * np;
  const double* Brow = Lb + (int64_t)j * NB * np;
  if (nk > 0) {
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < nk) {
        load_oper_tile(smem + (size_t)s * 2 * OPER_TILE, Arow, np, s * KB, tid);
        if (!is_diag) load_oper_tile(smem + (size_t)s * 2 * OPER_TILE + OPER_TILE, Brow, np, s * KB, tid);
      }
      cp_async_commit();
    }
    for (int ks = 0; ks < nk; ks++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        const int nx = ks + STAGES - 1;
        if (nx < nk) {
          const int s = nx % STAGES;
          load_oper_tile(smem + (size_t)s * 2 * OPER_TILE, Arow, np, nx * KB, tid);
          if (!is_diag) load_oper_tile(smem + (size_t)s * 2 * OPER_TILE + OPER_TILE, Brow, np, nx * KB, tid);
        }
        cp_async_commit();
      }
      const double* As = smem + (size_t)(ks % STAGES) * 2 * OPER_TILE;
      const double* Bs = is_diag ? As : (As + OPER_TILE);
#pragma unroll
      for (int k4 = 0; k4 < KB / 4; k4++) {
        double a[8], bf[4];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) a[mi] = As[(wm * 64 + mi * 8 + lr) * SA + k4 * 4 + lc];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = Bs[(wn * 32 + ni * 8 + lr) * SA + k4 * 4 + lc];
#pragma unroll
        for (int mi = 0; mi < 8; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) mma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---------------- phase B: C = AtA tile (damped) - acc, to shared memory ----------------
  double* Cs = smem;
  {
    const double* Ab = p.AtA + b * p.n * p.n;
    const double al = (p.alpha != nullptr) ? p.alpha[b] : 0.0;
    const double be = (p.beta != nullptr) ? p.beta[b] : 0.0;
#pragma unroll
    for (int mi = 0; mi < 8; mi++) {
      const int r = wm * 64 + mi * 8 + lr;
      const int64_t gr = (int64_t)i * NB + r;
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        const int c = wn * 32 + ni * 8 + lc * 2;
        double v[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int64_t gc = (int64_t)j * NB + c + u;
          double x;
          if (gr < p.n && gc < p.n) {
            x = Ab[gr * p.n + gc];
            if (gr == gc) x = x + (al * x + be);  // dense_solver.py:38-64 ; linear/utils.py:14-33
          } else {
            x = (gr == gc) ? 1.0 : 0.0;  // identity padding
          }
          v[u] = x - acc[mi][ni][u];
        }
        *reinterpret_cast<double2*>(&Cs[r * SC + c]) = make_double2(v[0], v[1]);
      }
    }
  }
  __syncthreads();

  double* Wj = p.W + ((int64_t)b * p.nblk + j) * NB * NB;
  int* flag = p.flags + b * p.nblk + j;

  if (is_diag) {
    // ---------------- phase C: potrf of the diagonal tile in shared memory ----------------
    __shared__ double colbuf[NB];
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    for (int c = 0; c < NB; c++) {
      __syncthreads();
      const double d = Cs[c * SC + c];
      if (tid == 0 && !(d > 0.0) && s_fail == 0) s_fail = j * NB + c + 1;
      const double sq = sqrt(d);
      const double inv = 1.0 / sq;
      if (tid >= c && tid < NB) {
        const double v = (tid == c) ? sq : Cs[tid * SC + c] * inv;
        colbuf[tid] = v;
        if (tid != c) Cs[tid * SC + c] = v;  // the pivot itself is still being read by other threads
      }
      __syncthreads();
      if (tid == 0) Cs[c * SC + c] = sq;
      // trailing update: 2 threads per row
      const int r = c + 1 + (tid >> 1);
      if (r < NB) {
        const double lrc = colbuf[r];
        for (int q = c + 1 + (tid & 1); q <= r; q += 2) Cs[r * SC + q] -= lrc * colbuf[q];
      }
    }
    __syncthreads();
    if (tid == 0 && s_fail != 0) atomicCAS(p.info + b, 0, s_fail);
    // store L_jj (lower; upper zeroed)
    for (int e = tid; e < NB * NB; e += CHOL_THREADS) {
      const int r = e >> 7, c = e & 127;
      Lb[((int64_t)j * NB + r) * np + (int64_t)j * NB + c] = (c <= r) ? Cs[r * SC + c] : 0.0;
    }
    // in-place inverse of the lower-triangular tile: column by column from the right
    for (int c = NB - 1; c >= 0; c--) {
      __syncthreads();
      const double x = 1.0 / Cs[c * SC + c];
      const int r = c + 1 + (tid >> 1);
      double s = 0.0;
      if (r < NB) {
        for (int k = c + 1 + (tid & 1); k <= r; k += 2) s += Cs[r * SC + k] * Cs[k * SC + c];
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      __syncthreads();
      if (r < NB && (tid & 1) == 0) Cs[r * SC + c] = -s * x;
      if (tid == 0) Cs[c * SC + c] = x;
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += CHOL_THREADS) {
      const int r = e >> 7, c = e & 127;
      Wj[e] = (c <= r) ? Cs[r * SC + c] : 0.0;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(flag), "r"(1) : "memory");
    }
    return;
  }

  // ---------------- phase D: L[i,j] = C W^T (DMMA, triangular k-range) ----------------
  if (tid == 0) {
    int v = 0;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(flag) : "memory");
      if (v == 0) __nanosleep(200);
    } while (v == 0);
  }
  __syncthreads();
  double* Ws = smem + NB * SC;
  double acc2[2][16][2];
#pragma unroll
  for (int mi = 0; mi < 2; mi++)
#pragma unroll
    for (int ni = 0; ni < 16; ni++) acc2[mi][ni][0] = acc2[mi][ni][1] = 0.0;
  constexpr int NKW = NB / KB;  // 8
#pragma unroll
  for (int s = 0; s < WSTAGES - 1; s++) {
    load_oper_tile(Ws + (size_t)s * OPER_TILE, Wj, NB, s * KB, tid);
    cp_async_commit();
  }
#pragma unroll
  for (int ks = 0; ks < NKW; ks++) {
    cp_async_wait<WSTAGES - 2>();
    __syncthreads();
    {
      const int nx = ks + WSTAGES - 1;
      if (nx < NKW) load_oper_tile(Ws + (size_t)(nx % WSTAGES) * OPER_TILE, Wj, NB, nx * KB, tid);
      cp_async_commit();
    }
    const double* Wst = Ws + (size_t)(ks % WSTAGES) * OPER_TILE;
#pragma unroll
    for (int k4 = 0; k4 < KB / 4; k4++) {
      const int kk = ks * KB + k4 * 4;
      double a[2];
#pragma unroll
      for (int mi = 0; mi < 2; mi++) a[mi] = Cs[(warp * 16 + mi * 8 + lr) * SC + kk + lc];
#pragma unroll
      for (int ni = 0; ni < 16; ni++) {
        if (ni * 8 + 7 >= kk) {  // W[c][k] == 0 for k > c: skip column blocks entirely above this k
          const double bfr = Wst[(ni * 8 + lr) * SA + k4 * 4 + lc];
          mma884(acc2[0][ni][0], acc2[0][ni][1], a[0], bfr);
          mma884(acc2[1][ni][0], acc2[1][ni][1], a[1], bfr);
        }
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mi = 0; mi < 2; mi++) {
    const int r = warp * 16 + mi * 8 + lr;
    double* dst = Lb + ((int64_t)i * NB + r) * np + (int64_t)j * NB;
#pragma unroll
    for (int ni = 0; ni < 16; ni++) {
      *reinterpret_cast<double2*>(dst + ni * 8 + lc * 2) = make_double2(acc2[mi][ni][0], acc2[mi][ni][1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Solve with the stored factor: forward  y_j = W_j (rhs_j - sum_{k<j} L[j,k] y_k),
//                               backward x_j = W_j^T (y_j - sum_{k>j} L[k,j]^T x_k).
struct SolveArgs {
  const double* L;
  const double* W;
  const double* rhs;  // [B,n]
  double* x;          // [B,n]
  int64_t B, n, np;
  int nblk;
};

constexpr int SOLVE_THREADS = 256;

__global__ void __launch_bounds__(SOLVE_THREADS, 2) chol_solve_kernel(SolveArgs p) {
  extern __shared__ __align__(16) double sm[];
  double* y = sm;               // [np]
  double* tmp = sm + p.np;      // [128]
  double* part = tmp + NB;      // [8][128]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x;
  const int64_t np = p.np;
  const double* Lb = p.L + b * np * np;
  const double* Wb = p.W + b * p.nblk * NB * NB;
  for (int64_t e = tid; e < np; e += SOLVE_THREADS) y[e] = (e < p.n) ? p.rhs[b * p.n + e] : 0.0;
  __syncthreads();
  // ---- forward ----
  for (int j = 0; j < p.nblk; j++) {
    const int K = j * NB;
    // each warp: 16 rows
    for (int rr = 0; rr < 16; rr += 2) {
      const int r0 = warp * 16 + rr;
      const double* row0 = Lb + ((int64_t)j * NB + r0) * np;
      const double* row1 = row0 + np;
      double s0 = 0.0, s1 = 0.0;
      for (int k = lane * 2; k < K; k += 64) {
        const double2 a0 = *reinterpret_cast<const double2*>(row0 + k);
        const double2 a1 = *reinterpret_cast<const double2*>(row1 + k);
        const double2 yy = *reinterpret_cast<const double2*>(y + k);
        s0 += a0.x * yy.x + a0.y * yy.y;
        s1 += a1.x * yy.x + a1.y * yy.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane == 0) {
        tmp[r0] = y[K + r0] - s0;
        tmp[r0 + 1] = y[K + r0 + 1] - s1;
      }
    }
    __syncthreads();
    const double* Wj = Wb + (int64_t)j * NB * NB;
    for (int rr = 0; rr < 16; rr++) {
      const int r = warp * 16 + rr;
      const double* wr = Wj + r * NB;
      double s = 0.0;
      for (int k = lane; k <= r; k += 32) s += wr[k] * tmp[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) y[K + r] = s;
    }
    __syncthreads();
  }
  // ---- backward ----
  for (int j = p.nblk - 1; j >= 0; j--) {
    const int K = j * NB;
    // part[w][c] = sum over rows r (this warp's share) of L[r][K+c] * x[r], r in [(j+1)*128, np)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int64_t r = (int64_t)(j + 1) * NB + warp; r < np; r += 8) {
      const double xr = y[r];
      const double* row = Lb + r * np + K + lane * 4;
      const double2 v0 = *reinterpret_cast<const double2*>(row);
      const double2 v1 = *reinterpret_cast<const double2*>(row + 2);
      a0 += v0.x * xr;
      a1 += v0.y * xr;
      a2 += v1.x * xr;
      a3 += v1.y * xr;
    }
    part[warp * NB + lane * 4 + 0] = a0;
    part[warp * NB + lane * 4 + 1] = a1;
    part[warp * NB + lane * 4 + 2] = a2;
    part[warp * NB + lane * 4 + 3] = a3;
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) s += part[w * NB + tid];
      tmp[tid] = y[K + tid] - s;
    }
    __syncthreads();
    // x_j[c] = sum_{r>=c} W[r][c] tmp[r]
    const double* Wj = Wb + (int64_t)j * NB * NB;
    a0 = a1 = a2 = a3 = 0.0;
    for (int r = warp; r < NB; r += 8) {
      const double tr = tmp[r];
      const double* row = Wj + r * NB + lane * 4;
      const double2 v0 = *reinterpret_cast<const double2*>(row);
      const double2 v1 = *reinterpret_cast<const double2*>(row + 2);
      a0 += v0.x * tr;
      a1 += v0.y * tr;
      a2 += v1.x * tr;
      a3 += v1.y * tr;
    }
    part[warp * NB + lane * 4 + 0] = a0;
    part[warp * NB + lane * 4 + 1] = a1;
    part[warp * NB + lane * 4 + 2] = a2;
    part[warp * NB + lane * 4 + 3] = a3;
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) s += part[w * NB + tid];
      y[K + tid] = s;
    }
    __syncthreads();
  }
  for (int64_t e = tid; e < p.n; e += SOLVE_THREADS) p.x[b * p.n + e] = y[e];
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace thb

extern "C" {

int64_t thb_potrf_workspace_bytes(int64_t B, int64_t n) {
  if (B <= 0 || n <= 0) return 0;
  const int64_t nblk = (n + thb::NB - 1) / thb::NB, np = nblk * thb::NB;
  int64_t bytes = B * np * np * 8;                              // L
  bytes += B * nblk * thb::NB * thb::NB * 8;                    // W
  bytes += thb::align_up(B * nblk * 4, 256);                    // flags
  return bytes;
}

int thb_potrf_potrs_f64(const double* AtA, const double* rhs, const double* alpha, const double* beta, double* x, int32_t* info,
                        int64_t B, int64_t n, void* workspace, int64_t workspace_bytes, thb_stream_t stream) {
  if (B < 0 || n < 0 || AtA == nullptr || rhs == nullptr || x == nullptr || info == nullptr || workspace == nullptr)
    return THB_ERR_BAD_ARG;
  if (B == 0 || n == 0) return THB_OK;
  if (workspace_bytes < thb_potrf_workspace_bytes(B, n)) return THB_ERR_BAD_ARG;
  cudaStream_t cs = thb_cs(stream);
  const int nblk = (int)((n + thb::NB - 1) / thb::NB);
  const int64_t np = (int64_t)nblk * thb::NB;
  char* ws = static_cast<char*>(workspace);
  double* L = reinterpret_cast<double*>(ws);
  double* W = L + B * np * np;
  int* flags = reinterpret_cast<int*>(W + B * nblk * thb::NB * thb::NB);
  THB_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * nblk * 4, cs));
  THB_CUDA(cudaMemsetAsync(info, 0, (size_t)B * 4, cs));
  static bool attr_set = false;
  if (!attr_set) {
    THB_CUDA(cudaFuncSetAttribute(thb::chol_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thb::CHOL_SMEM));
    attr_set = true;
  }
  thb::CholArgs a;
  a.AtA = AtA; a.alpha = alpha; a.beta = beta; a.L = L; a.W = W; a.flags = flags; a.info = info;
  a.B = B; a.n = n; a.np = np; a.nblk = nblk;
  for (int j = 0; j < nblk; j++) {
    a.j = j;
    const int64_t grid = (int64_t)(nblk - j) * B;
    thb::chol_col_kernel<<<(unsigned)grid, thb::CHOL_THREADS, thb::CHOL_SMEM, cs>>>(a);
    THB_CHECK_LAUNCH();
  }
  thb::SolveArgs s;
  s.L = L; s.W = W; s.rhs = rhs; s.x = x; s.B = B; s.n = n; s.np = np; s.nblk = nblk;
  const size_t ssm = (size_t)(np + thb::NB + 8 * thb::NB) * 8;
  static size_t solve_smem_set = 0;
  if (ssm > 48 * 1024 && ssm > solve_smem_set) {
    THB_CUDA(cudaFuncSetAttribute(thb::chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
    solve_smem_set = ssm;
  }
  thb::chol_solve_kernel<<<(unsigned)B, thb::SOLVE_THREADS, ssm, cs>>>(s);
  THB_CHECK_LAUNCH();
  return THB_OK;
}

}  // extern "C"
