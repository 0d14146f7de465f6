// Micro-benchmark (B200): latency of dependent FP64 chains -- DFMA, DADD, LDS->DFMA, 64-bit SHFL+DADD -- with 1..16 warps per SM,
// and the same work split over 1 / 2 / 4 independent accumulators.  Prints cycles per chain step.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int n, double* out, long long* cyc, double a0) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double a = a0, b = 1.0000001, c = 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < n; i++) a = fma(a, b, c);
  if (mode == 1) for (int i = 0; i < n; i++) a = a + b;
  if (mode == 2) for (int i = 0; i < n; i++) a = fma(sm[(i * 33 + threadIdx.x) & 4095], b, a);
  if (mode == 3) for (int i = 0; i < n; i++) a += __shfl_xor_sync(0xffffffffu, a, 1);
  if (mode == 4) for (int i = 0; i < n; i += 2) { a = fma(a, b, c); a1 = fma(a1, b, c); }
  if (mode == 5) for (int i = 0; i < n; i += 4) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
  if (mode == 6) { float f = (float)a0, g = 1.0000001f, h = 1e-9f; for (int i = 0; i < n; i++) f = fmaf(f, g, h); a = f; }
  if (mode == 7) for (int i = 0; i < n; i += 4) { a = fma(sm[(i * 33 + threadIdx.x) & 4095], b, a); a1 = fma(sm[(i * 33 + 33 + threadIdx.x) & 4095], b, a1);
                                                  a2 = fma(sm[(i * 33 + 66 + threadIdx.x) & 4095], b, a2); a3 = fma(sm[(i * 33 + 99 + threadIdx.x) & 4095], b, a3); }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a1 + a2 + a3;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA chain", "DADD chain", "LDS->DFMA chain", "SHFL64+DADD chain", "DFMA 2 chains", "DFMA 4 chains", "FFMA chain", "LDS->DFMA 4 chains"};
  const int n = 4096;
  for (int mode = 0; mode < 8; mode++)
    for (int thr : {32, 128, 512, 1024}) {
      for (int rep = 0; rep < 2; rep++) k<<<148, thr>>>(mode, n, out, cyc, 1.0);
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-20s threads/SM %4d : %.1f cycles per step\n", names[mode], thr, (double)h / n);
    }
  return 0;
}
