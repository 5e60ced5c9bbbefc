// FFMA vs FFMA2 (packed fp32x2) issue rate on sm_100: 16 independent accumulator chains per thread, 8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float acc[16];
  f32x2 acc2[16];
  for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x + i; acc2[i] = (unsigned long long)(threadIdx.x + i) * 0x100000001ull; }
  f32x2 a2, b2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(b), "f"(b));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PACKED) acc2[i] = fma2(acc2[i], a2, b2);
      else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += PACKED ? __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32)) : acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
  const int iters = 20000;
  for (int p = 0; p < 2; ++p) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (p) k<1><<<148 * 4, 256>>>(out, iters, 1.0001f, 0.5f); else k<0><<<148 * 4, 256>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * 4 * 256 * 16.0 * iters * (p ? 2 : 1);
    printf("%s: %.3f ms, %.1f TFLOP/s fp32 (%.1f FMA/clk/SM at 1.9 GHz)\n", p ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
  }
  return 0;
}
