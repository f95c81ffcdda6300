// Multifrontal (supernodal) batched block-sparse Cholesky, fp64, sm_100a -- numeric phase of layout "front".
//
// Replaces BaSpaCho's batched factor / solve behind NumericDecomposition::factor / solve
// (theseus/extlib/baspacho_solver_cuda.cu:203-214, 282-287) for B problems that share ONE structure.  Symbolic side:
// theseus_b200/frontal.py (fronts, depth schedule, relative maps); data layout: include/thb200.h (thb_front_plan).
//
// A front t is a dense symmetric matrix  F_t = [[D, .], [P, C]]  (w pivots, b border rows).  Per depth, deepest first:
//   assemble   F_t = panel of AtA entries (+ LM damping on the diagonal) + sum over children c of  extend-add(C_c)
//   factor     D = L L^T,  P <- P L^-T,  C <- C - P P^T       (C_t goes to the parent through the update-matrix arena)
// Small fronts (r <= 160): ONE CTA per (front, item), the whole front in shared memory: scatter-add of the children, scalar
//   right-looking pivots on the r x w panel, the rank-w update of C on the FP64 tensor pipe (mma.sync m8n8k4 DMMA, 16x16 macro
//   tiles per warp) -- every byte of a small front is read once and written once.
// Big fronts: assembled into a padded dense matrix in global memory (row tiles in shared memory, children added in a fixed order),
//   then the DMMA dense kernel of thb_chol_dense.cu in partial mode (left-looking tiles: the Schur complement is accumulated in
//   registers over all pivot columns and written once), then the panel is copied to the factor storage.
// Substitutions: per front one CTA per item, chunked by 32 pivot columns: a 32 x 32 triangular block is solved by one warp with
//   shuffles, the rest of the panel is a row-contiguous mat-vec; the forward pass hands border vectors to the parent through a
//   second ping-pong arena, the backward pass gathers the ancestors' solution.  The permutation is folded into the first load /
//   last store.  No atomics on data anywhere: bitwise reproducible, independent of B.
#include "thb_common.cuh"

namespace thb {

#ifdef THB_SIMT_EMU
// host emulation (tests/simt): the m8n8k4 fragment semantics with shuffles
__device__ inline void front_mma884(double& c0, double& c1, double a, double b) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
  for (int k = 0; k < 4; k++) {
    const double av = __shfl_sync(0xffffffffu, a, lr * 4 + k);
    const double b0 = __shfl_sync(0xffffffffu, b, (2 * lc) * 4 + k);
    const double b1 = __shfl_sync(0xffffffffu, b, (2 * lc + 1) * 4 + k);
    c0 += av * b0;
    c1 += av * b1;
  }
}
#else
__device__ __forceinline__ void front_mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// smallest v >= x with v % 16 == 4: row stride (doubles) with conflict-free DMMA fragment loads / accumulator stores
__host__ __device__ __forceinline__ int front_pad_ld(int x) { return ((x + 11) / 16) * 16 + 4; }

struct FrontArgs {
  thb_front_plan p;
  int s0;                   // first entry of this launch in p.sched
  int64_t B;
  double* factor;           // [B, data_size]
  const double* alpha;      // [B] or null
  const double* beta;       // [B] or null
  double* arena_cur;        // [B, arena_size] of this depth's parity
  const double* arena_child;
  int32_t* info;
};

// ------------------------------------------------------------------------------------------------ small fronts
template <int THREADS>
__global__ void __launch_bounds__(THREADS) front_small_kernel(FrontArgs a) {
  extern __shared__ double sm[];
  const thb_front_plan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int64_t item = blockIdx.x;
  const int t = p.sched[a.s0 + blockIdx.y];
  const int w = p.f_w[t], b = p.f_b[t], r = w + b;
  const int b16 = (b + 15) & ~15, w4 = (w + 3) & ~3;
  const int ldp = front_pad_ld(w4), ldc = front_pad_ld(b16);
  const int prow = w + b16;
  double* PN = sm;                 // [prow][ldp]  panel: pivot block on top, border rows below, zero padding
  double* CB = sm + prow * ldp;    // [b16][ldc]   update matrix (lower triangle meaningful)
  for (int e = tid; e < prow * ldp + b16 * ldc; e += THREADS) sm[e] = 0.0;
  __syncthreads();
  double* Lg = a.factor + item * p.data_size + p.f_panel_off[t];
  {
    const double al = a.alpha != nullptr ? a.alpha[item] : 0.0;
    const double be = a.beta != nullptr ? a.beta[item] : 0.0;
    for (int e = tid; e < r * w; e += THREADS) {
      const int i = e / w, j = e - i * w;
      double v = Lg[e];
      if (i == j) v = v + (al * v + be);   // linear/utils.py:14-33: diag <- diag (1 + alpha) + beta
      PN[i * ldp + j] = v;
    }
  }
  __syncthreads();
  // ---- extend-add of the children's update matrices, one child after the other (fixed order) ----
  for (int ci = p.child_ptr[t]; ci < p.child_ptr[t + 1]; ci++) {
    const int c = p.child_list[ci];
    const int bc = p.f_b[c], ldg = p.f_cb_ld[c];
    const double* src = a.arena_child + item * p.arena_size + p.f_cb_off[c];
    const int32_t* rel = p.f_rel + p.rel_ptr[c];
    for (int i = warp; i < bc; i += NW) {
      const int ri = rel[i];
      const double* srow = src + (int64_t)i * ldg;
      for (int j = lane; j <= i; j += 32) {
        const int rj = rel[j];
        const double v = srow[j];
        if (rj < w) PN[ri * ldp + rj] += v;
        else CB[(ri - w) * ldc + (rj - w)] += v;
      }
    }
    __syncthreads();
  }
  // ---- pivots: right-looking on the r x w panel (columns k+1..w-1 of all rows; the rank-w update of C comes after) ----
  for (int k = 0; k < w; k++) {
    const double pkk = PN[k * ldp + k];
    double s;
    if (pkk > 0.0) {
      s = 1.0 / sqrt(pkk);
    } else {
      s = 1.0;
      if (tid == 0) atomicCAS(a.info + item, 0, p.f_first[t] + k + 1);
    }
    for (int i = k + 1 + tid; i < r; i += THREADS) PN[i * ldp + k] *= s;
    __syncthreads();
    if (tid == 0) PN[k * ldp + k] = pkk * s;
    const int nc = w - k - 1;
    if (nc > 0) {
      for (int e = tid; e < nc * (r - k - 1); e += THREADS) {
        const int i = k + 1 + e / nc, j = k + 1 + e % nc;
        if (i >= j) PN[i * ldp + j] -= PN[i * ldp + k] * PN[j * ldp + k];
      }
    }
    __syncthreads();
  }
  // ---- C -= P P^T on the FP64 tensor pipe: 16 x 16 macro tiles of the lower triangle, one warp each ----
  if (b > 0) {
    const int lr = lane >> 2, lc = lane & 3;
    const int nt = b16 / 16, ntl = nt * (nt + 1) / 2;
    const double* P = PN + w * ldp;
    for (int q = warp; q < ntl; q += NW) {
      int ti = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= q) ti++;
      while (ti * (ti + 1) / 2 > q) ti--;
      const int tj = q - ti * (ti + 1) / 2;
      double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
      const double* Pa = P + (16 * ti + lr) * ldp + lc;
      const double* Pb = P + (16 * tj + lr) * ldp + lc;
      for (int k4 = 0; k4 < w4; k4 += 4) {
        const double a0 = Pa[k4], a1 = Pa[8 * ldp + k4];
        const double b0 = Pb[k4], b1 = Pb[8 * ldp + k4];
        front_mma884(acc[0][0][0], acc[0][0][1], a0, b0);
        front_mma884(acc[0][1][0], acc[0][1][1], a0, b1);
        front_mma884(acc[1][0][0], acc[1][0][1], a1, b0);
        front_mma884(acc[1][1][0], acc[1][1][1], a1, b1);
      }
#pragma unroll
      for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++) {
          double* d = CB + (16 * ti + 8 * mi + lr) * ldc + 16 * tj + 8 * ni + 2 * lc;
          d[0] -= acc[mi][ni][0];
          d[1] -= acc[mi][ni][1];
        }
    }
  }
  __syncthreads();
  // ---- write back: the panel (zeros above the diagonal of the pivot block), the update matrix (lower triangle) ----
  for (int e = tid; e < r * w; e += THREADS) {
    const int i = e / w, j = e - i * w;
    Lg[e] = (j > i) ? 0.0 : PN[i * ldp + j];
  }
  if (b > 0) {
    double* dst = a.arena_cur + item * p.arena_size + p.f_cb_off[t];
    const int ldg = p.f_cb_ld[t];
    for (int i = warp; i < b; i += NW)
      for (int j = lane; j <= i; j += 32) dst[(int64_t)i * ldg + j] = CB[i * ldc + j];
  }
}

// ------------------------------------------------------------------------------------------------ big fronts: assemble / extract
constexpr int ASM_ROWS = 16, ASM_THREADS = 256;

__device__ __forceinline__ int front_map_big(int l, int w, int wpad) { return l < w ? l : wpad + (l - w); }

// grid: x = item, y = row tile (np / ASM_ROWS).  Builds rows [R0, R0 + ASM_ROWS) of the padded front matrix F (np x np, row-major):
// zeros, identity on the padding, the panel's AtA entries (+ damping), the children's update matrices; written once.
__global__ void __launch_bounds__(ASM_THREADS) front_assemble_kernel(FrontArgs a, int t) {
  extern __shared__ double sm[];
  const thb_front_plan& p = a.p;
  const int tid = threadIdx.x;
  const int64_t item = blockIdx.x;
  const int w = p.f_w[t], b = p.f_b[t], r = w + b, wpad = p.f_wpad[t], np = p.f_np[t];
  const int R0 = blockIdx.y * ASM_ROWS;
  double* buf = sm;   // [ASM_ROWS][np]
  for (int e = tid; e < ASM_ROWS * np; e += ASM_THREADS) buf[e] = 0.0;
  __syncthreads();
  const double* Lg = a.factor + item * p.data_size + p.f_panel_off[t];
  const double al = a.alpha != nullptr ? a.alpha[item] : 0.0;
  const double be = a.beta != nullptr ? a.beta[item] : 0.0;
  for (int e = tid; e < ASM_ROWS * wpad; e += ASM_THREADS) {
    const int rr = e / wpad, j = e - rr * wpad;
    const int fr = R0 + rr;
    int i = -1;   // local front row
    if (fr < w) i = fr;
    else if (fr >= wpad && fr - wpad + w < r) i = fr - wpad + w;
    if (i >= 0) {
      if (j < w && !(i < w && j > i)) {
        double v = Lg[(int64_t)i * w + j];
        if (i == j) v = v + (al * v + be);
        buf[rr * np + j] = v;
      }
    } else if (j == fr) {
      buf[rr * np + j] = 1.0;   // padding inside the pivot columns
    }
  }
  for (int rr = tid; rr < ASM_ROWS; rr += ASM_THREADS) {
    const int fr = R0 + rr;
    if (fr >= wpad + b) buf[rr * np + fr] = 1.0;   // padding after the border rows
  }
  __syncthreads();
  for (int ci = p.child_ptr[t]; ci < p.child_ptr[t + 1]; ci++) {
    const int c = p.child_list[ci];
    const int bc = p.f_b[c], ldg = p.f_cb_ld[c];
    const int32_t* rel = p.f_rel + p.rel_ptr[c];
    // child rows whose image lies in this row tile: [i0, i1)  (rel is increasing)
    int lo = 0, hi = bc;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (front_map_big(rel[mid], w, wpad) < R0) lo = mid + 1; else hi = mid; }
    const int i0 = lo;
    hi = bc;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (front_map_big(rel[mid], w, wpad) < R0 + ASM_ROWS) lo = mid + 1; else hi = mid; }
    const int i1 = lo;
    if (i1 > i0) {
      const double* src = a.arena_child + item * p.arena_size + p.f_cb_off[c];
      const int nrow = i1 - i0;
      for (int e = tid; e < nrow * bc; e += ASM_THREADS) {
        const int i = i0 + e / bc, j = e % bc;
        if (j <= i) buf[(front_map_big(rel[i], w, wpad) - R0) * np + front_map_big(rel[j], w, wpad)] += src[(int64_t)i * ldg + j];
      }
    }
    __syncthreads();
  }
  double* F = a.arena_cur + item * p.arena_size + p.f_fr_off[t] + (int64_t)R0 * np;
  for (int e = tid; e < ASM_ROWS * np; e += ASM_THREADS) F[e] = buf[e];
}

// grid: x = item, y = chunk.  factor panel <- the pivot columns of the factored front matrix
__global__ void __launch_bounds__(256) front_extract_kernel(FrontArgs a, int t) {
  const thb_front_plan& p = a.p;
  const int64_t item = blockIdx.x;
  const int w = p.f_w[t], b = p.f_b[t], r = w + b, wpad = p.f_wpad[t], np = p.f_np[t];
  const double* F = a.arena_cur + item * p.arena_size + p.f_fr_off[t];
  double* Lg = a.factor + item * p.data_size + p.f_panel_off[t];
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < r * w; e += gridDim.y * blockDim.x) {
    const int i = e / w, j = e - i * w;
    const int fr = i < w ? i : wpad + (i - w);
    Lg[e] = (i < w && j > i) ? 0.0 : F[(int64_t)fr * np + j];
  }
}

// ------------------------------------------------------------------------------------------------ substitutions
struct FrontSolveArgs {
  thb_front_plan p;
  int s0;
  int64_t B;
  const double* factor;
  const double* rhs;        // [B, n] original order (forward)
  double* x;                // [B, n] original order (backward)
  double* work;             // [B, n] permuted: y after the forward pass, x after the backward pass
  double* v_cur;            // border vectors of this depth's parity
  const double* v_child;
};

// One warp: solve T y = u (lower triangular cw x cw, row stride 33) -- lane i owns u_i
__device__ __forceinline__ double front_warp_trsv_lower(const double* T, double ui, int cw, int lane) {
  const double rd = lane < cw ? 1.0 / T[lane * 33 + lane] : 0.0;
  for (int k = 0; k < cw; k++) {
    const double yk = __shfl_sync(0xffffffffu, ui * rd, k);
    if (lane == k) ui = yk;
    else if (lane > k && lane < cw) ui -= T[lane * 33 + k] * yk;
  }
  return ui;
}
// One warp: solve T^T x = t
__device__ __forceinline__ double front_warp_trsv_upper(const double* T, double ti, int cw, int lane) {
  const double rd = lane < cw ? 1.0 / T[lane * 33 + lane] : 0.0;
  for (int k = cw - 1; k >= 0; k--) {
    const double xk = __shfl_sync(0xffffffffu, ti * rd, k);
    if (lane == k) ti = xk;
    else if (lane < k) ti -= T[k * 33 + lane] * xk;
  }
  return ti;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) front_forward_kernel(FrontSolveArgs a) {
  extern __shared__ double sm[];
  const thb_front_plan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t item = blockIdx.x;
  const int t = p.sched[a.s0 + blockIdx.y];
  const int w = p.f_w[t], b = p.f_b[t], r = w + b, first = p.f_first[t];
  double* u = sm;            // [r]
  double* T = sm + ((r + 1) & ~1);   // [32][33]
  const double* Lg = a.factor + item * p.data_size + p.f_panel_off[t];
  for (int i = tid; i < r; i += THREADS) u[i] = i < w ? a.rhs[item * p.n + p.perm[first + i]] : 0.0;
  __syncthreads();
  for (int ci = p.child_ptr[t]; ci < p.child_ptr[t + 1]; ci++) {
    const int c = p.child_list[ci];
    const int bc = p.f_b[c];
    const double* uc = a.v_child + item * p.varena_size + p.f_u_off[c];
    const int32_t* rel = p.f_rel + p.rel_ptr[c];
    for (int i = tid; i < bc; i += THREADS) u[rel[i]] += uc[i];
    __syncthreads();
  }
  for (int k0 = 0; k0 < w; k0 += 32) {
    const int cw = min(32, w - k0);
    for (int e = tid; e < cw * cw; e += THREADS) {
      const int i = e / cw, j = e - i * cw;
      T[i * 33 + j] = Lg[(int64_t)(k0 + i) * w + k0 + j];
    }
    __syncthreads();
    if (warp == 0) {
      const double ui = front_warp_trsv_lower(T, lane < cw ? u[k0 + lane] : 0.0, cw, lane);
      if (lane < cw) u[k0 + lane] = ui;
    }
    __syncthreads();
    for (int i = k0 + cw + tid; i < r; i += THREADS) {
      const double* row = Lg + (int64_t)i * w + k0;
      double s = 0.0;
      for (int k = 0; k < cw; k++) s += row[k] * u[k0 + k];
      u[i] -= s;
    }
    __syncthreads();
  }
  for (int i = tid; i < w; i += THREADS) a.work[item * p.n + first + i] = u[i];
  if (b > 0) {
    double* ub = a.v_cur + item * p.varena_size + p.f_u_off[t];
    for (int i = tid; i < b; i += THREADS) ub[i] = u[w + i];
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) front_backward_kernel(FrontSolveArgs a) {
  extern __shared__ double sm[];
  const thb_front_plan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t item = blockIdx.x;
  const int t = p.sched[a.s0 + blockIdx.y];
  const int w = p.f_w[t], b = p.f_b[t], r = w + b, first = p.f_first[t];
  double* xf = sm;                       // [r] pivots (y, then x) followed by the border rows' x
  double* T = sm + ((r + 1) & ~1);       // [32][33]
  double* part = T + 32 * 33 + 1;        // [THREADS]
  const double* Lg = a.factor + item * p.data_size + p.f_panel_off[t];
  double* wk = a.work + item * p.n;
  const int32_t* rows = p.f_rows + p.rows_ptr[t];
  for (int i = tid; i < r; i += THREADS) xf[i] = i < w ? wk[first + i] : wk[rows[i - w]];
  __syncthreads();
  const int nchunk = (w + 31) / 32;
  for (int ch = nchunk - 1; ch >= 0; ch--) {
    const int k0 = ch * 32, cw = min(32, w - k0);
    // t_k = y_k - sum_{i >= k0 + cw} L[i][k0 + k] x_i : cw consecutive threads per row, THREADS / cw rows per pass
    const int RG = THREADS / cw;
    const int kk = tid % cw, rg = tid / cw;
    double acc = 0.0;
    if (rg < RG)
      for (int i = k0 + cw + rg; i < r; i += RG) acc += Lg[(int64_t)i * w + k0 + kk] * xf[i];
    part[tid] = acc;
    for (int e = tid; e < cw * cw; e += THREADS) {
      const int i = e / cw, j = e - i * cw;
      T[i * 33 + j] = Lg[(int64_t)(k0 + i) * w + k0 + j];
    }
    __syncthreads();
    if (warp == 0) {
      double tk = 0.0;
      if (lane < cw) {
        tk = xf[k0 + lane];
        for (int g = 0; g < RG; g++) tk -= part[g * cw + lane];
      }
      tk = front_warp_trsv_upper(T, tk, cw, lane);
      if (lane < cw) xf[k0 + lane] = tk;
    }
    __syncthreads();
  }
  for (int i = tid; i < w; i += THREADS) {
    const double v = xf[i];
    wk[first + i] = v;
    a.x[item * p.n + p.perm[first + i]] = v;
  }
}

static inline int front_threads_of_class(int cls) { return cls == 0 ? 64 : (cls == 1 ? 128 : 256); }

template <typename K>
static inline int front_set_smem(K kernel, size_t bytes, size_t* cache) {
  if (bytes > 48 * 1024 && bytes > *cache) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    *cache = bytes;
  }
  return 0;
}

}  // namespace thb

extern "C" {

int64_t thb_front_small_smem_bytes(int32_t w, int32_t b) {
  const int b16 = (b + 15) & ~15, w4 = (w + 3) & ~3;
  return (int64_t)((w + b16) * thb::front_pad_ld(w4) + b16 * thb::front_pad_ld(b16)) * 8;
}

int thb_front_factor_f64(const thb_front_plan* p, const int64_t* launches, int64_t num_launches, double* factor, const double* alpha,
                         const double* beta, double* arena, void* dense_ws, int64_t dense_ws_bytes, int32_t* info, int64_t B,
                         thb_stream_t stream) {
  if (p == nullptr || launches == nullptr || factor == nullptr || arena == nullptr || info == nullptr || B < 0) return THB_ERR_BAD_ARG;
  if (B == 0 || p->S == 0) return THB_OK;
  if (B > 65535LL * 32768LL) return THB_ERR_UNSUPPORTED;
  cudaStream_t cs = thb_cs(stream);
  THB_CUDA(cudaMemsetAsync(info, 0, (size_t)B * 4, cs));
  static size_t smem_set[3] = {0, 0, 0}, asm_set = 0;
  for (int64_t l = 0; l < num_launches; l++) {
    const int64_t* L = launches + l * THB_FRONT_LAUNCH_COLS;
    const int depth = (int)L[0], cls = (int)L[1], begin = (int)L[2], count = (int)L[3];
    thb::FrontArgs a;
    a.p = *p; a.s0 = begin; a.B = B; a.factor = factor; a.alpha = alpha; a.beta = beta; a.info = info;
    a.arena_cur = arena + (int64_t)(depth & 1) * B * p->arena_size;
    a.arena_child = arena + (int64_t)((depth + 1) & 1) * B * p->arena_size;
    if (cls < 3) {
      const size_t smem = (size_t)L[4];
#ifndef THB_SIMT_EMU
      if (smem > 227 * 1024) return THB_ERR_UNSUPPORTED;
#endif
      const dim3 grid((unsigned)B, (unsigned)count);
      if (count > 65535) return THB_ERR_UNSUPPORTED;
      if (cls == 0) {
        int rc = thb::front_set_smem(thb::front_small_kernel<64>, smem, &smem_set[0]); if (rc) return rc;
        thb::front_small_kernel<64><<<grid, 64, smem, cs>>>(a);
      } else if (cls == 1) {
        int rc = thb::front_set_smem(thb::front_small_kernel<128>, smem, &smem_set[1]); if (rc) return rc;
        thb::front_small_kernel<128><<<grid, 128, smem, cs>>>(a);
      } else {
        int rc = thb::front_set_smem(thb::front_small_kernel<256>, smem, &smem_set[2]); if (rc) return rc;
        thb::front_small_kernel<256><<<grid, 256, smem, cs>>>(a);
      }
      THB_CHECK_LAUNCH();
    } else {
      const int64_t np = L[5], nb_piv = L[6], fr_off = L[7], first = L[8];
      const int t = (int)L[9];
#ifdef THB_SIMT_EMU
      (void)np; (void)nb_piv; (void)fr_off; (void)first; (void)t; (void)dense_ws; (void)dense_ws_bytes; (void)asm_set;
      return THB_ERR_UNSUPPORTED;   // the DMMA dense kernel is not part of the host emulation
#else
      if (dense_ws == nullptr || np % 128 != 0 || np > 8192) return THB_ERR_BAD_ARG;
      const size_t smem = (size_t)thb::ASM_ROWS * np * 8;
      int rc = thb::front_set_smem(thb::front_assemble_kernel, smem, &asm_set); if (rc) return rc;
      thb::front_assemble_kernel<<<dim3((unsigned)B, (unsigned)(np / thb::ASM_ROWS)), thb::ASM_THREADS, smem, cs>>>(a, t);
      THB_CHECK_LAUNCH();
      rc = thb_potrf_partial_inplace_f64(a.arena_cur + fr_off, p->arena_size, np, (int32_t)nb_piv, (int32_t)first, info, B, dense_ws,
                                         dense_ws_bytes, stream);
      if (rc != THB_OK) return rc;
      thb::front_extract_kernel<<<dim3((unsigned)B, 8), 256, 0, cs>>>(a, t);
      THB_CHECK_LAUNCH();
#endif
    }
  }
  return THB_OK;
}

int thb_front_solve_f64(const thb_front_plan* p, const int64_t* launches, int64_t num_launches, const double* factor, const double* rhs,
                        double* x, double* work, double* varena, int64_t B, thb_stream_t stream) {
  if (p == nullptr || launches == nullptr || factor == nullptr || rhs == nullptr || x == nullptr || work == nullptr || varena == nullptr ||
      B < 0)
    return THB_ERR_BAD_ARG;
  if (B == 0 || p->S == 0) return THB_OK;
  cudaStream_t cs = thb_cs(stream);
  static size_t fw_set[3] = {0, 0, 0}, bw_set[3] = {0, 0, 0};
  for (int pass = 0; pass < 2; pass++) {
    for (int64_t q = 0; q < num_launches; q++) {
      const int64_t l = pass == 0 ? q : num_launches - 1 - q;   // forward: deepest first; backward: roots first
      const int64_t* L = launches + l * THB_FRONT_LAUNCH_COLS;
      const int depth = (int)L[0], cls = (int)L[1], begin = (int)L[2], count = (int)L[3];
      thb::FrontSolveArgs a;
      a.p = *p; a.s0 = begin; a.B = B; a.factor = factor; a.rhs = rhs; a.x = x; a.work = work;
      a.v_cur = varena + (int64_t)(depth & 1) * B * p->varena_size;
      a.v_child = varena + (int64_t)((depth + 1) & 1) * B * p->varena_size;
      const int kc = cls > 2 ? 2 : cls;
      const int threads = thb::front_threads_of_class(kc);
      const int64_t r_max = cls > 2 ? L[5] : (cls == 0 ? 48 : (cls == 1 ? 96 : 160));   // class-3 launches carry np >= r
      const size_t smem = (size_t)(((r_max + 1) & ~1LL) + 32 * 33 + 1 + (pass == 1 ? threads : 0) + 2) * 8;
      const dim3 grid((unsigned)B, (unsigned)count);
      if (pass == 0) {
        if (kc == 0) { int rc = thb::front_set_smem(thb::front_forward_kernel<64>, smem, &fw_set[0]); if (rc) return rc;
                       thb::front_forward_kernel<64><<<grid, 64, smem, cs>>>(a); }
        else if (kc == 1) { int rc = thb::front_set_smem(thb::front_forward_kernel<128>, smem, &fw_set[1]); if (rc) return rc;
                            thb::front_forward_kernel<128><<<grid, 128, smem, cs>>>(a); }
        else { int rc = thb::front_set_smem(thb::front_forward_kernel<256>, smem, &fw_set[2]); if (rc) return rc;
               thb::front_forward_kernel<256><<<grid, 256, smem, cs>>>(a); }
      } else {
        if (kc == 0) { int rc = thb::front_set_smem(thb::front_backward_kernel<64>, smem, &bw_set[0]); if (rc) return rc;
                       thb::front_backward_kernel<64><<<grid, 64, smem, cs>>>(a); }
        else if (kc == 1) { int rc = thb::front_set_smem(thb::front_backward_kernel<128>, smem, &bw_set[1]); if (rc) return rc;
                            thb::front_backward_kernel<128><<<grid, 128, smem, cs>>>(a); }
        else { int rc = thb::front_set_smem(thb::front_backward_kernel<256>, smem, &bw_set[2]); if (rc) return rc;
               thb::front_backward_kernel<256><<<grid, 256, smem, cs>>>(a); }
      }
      THB_CHECK_LAUNCH();
    }
  }
  return THB_OK;
}

}  // extern "C"
