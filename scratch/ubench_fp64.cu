// Microbenchmark: FP64 DFMA vs DMMA (mma.sync f64) throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  double x = threadIdx.x * 1e-9 + 1.0, y = 1e-9;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma884_kernel(double* out, int iters) {
  double c[8][2];
  double a = threadIdx.x * 1e-9 + 1.0, b = 1e-9 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = i + 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef HAVE_M16N8K8
__global__ void dmma1688_kernel(double* out, int iters) {
  double c[4][4];
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-9 + i;
  b[0] = 1e-9; b[1] = 2e-9;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif
#ifdef HAVE_M16N8K16
__global__ void dmma16816_kernel(double* out, int iters) {
  double c[4][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1e-9 * i;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
  int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32; int blocks = 148 * (warps <= 16 ? 2 : 1);
    float ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    double flops = 2.0 * 16 * iters * (double)threads * blocks;
    printf("DFMA  warps/blk %2d blocks %d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms * 1e-9);
    ms = timeit([&] { dmma884_kernel<<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * blocks;
    printf("DMMA884 warps/blk %2d blocks %d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms * 1e-9);
#ifdef HAVE_M16N8K8
    ms = timeit([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 16 * 8 * 8 * 4 * iters * (double)(threads / 32) * blocks;
    printf("DMMA1688 warps/blk %2d blocks %d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms * 1e-9);
#endif
#ifdef HAVE_M16N8K16
    ms = timeit([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters); });
    flops = 2.0 * 16 * 8 * 16 * 4 * iters * (double)(threads / 32) * blocks;
    printf("DMMA16816 warps/blk %2d blocks %d: %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, flops / ms * 1e-9);
#endif
  }
  CK(cudaGetLastError());
  return 0;
}
