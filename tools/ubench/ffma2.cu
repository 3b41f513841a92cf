// Micro-benchmark (development aid): issue rate and dependent latency of the packed fp32 instructions of sm_100
// (fma.rn.f32x2 -> FFMA2) against scalar FFMA, per SM, as a function of the number of warps.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 tools/ubench/ffma2.cu && /tmp/ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* cycles, float* sink) {
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {  // FFMA: 16 independent chains
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (float)(threadIdx.x + j) * 1e-3f;
    const float a = 1.0001f, b = 1e-4f * (float)iters;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += v[j];
  } else if (MODE == 1) {  // FFMA2: 16 independent chains of pairs
    uint64_t v[16];
    uint64_t a, b;
    { const float af = 1.0001f, bf = 1e-4f * (float)iters; asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af)); asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf)); }
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float f = (float)(threadIdx.x + j) * 1e-3f; asm("mov.b64 %0, {%1, %1};" : "=l"(v[j]) : "f"(f)); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(a), "l"(b));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[j])); acc += lo + hi; }
  } else if (MODE == 2) {  // FFMA: one dependent chain (latency)
    float v = (float)threadIdx.x * 1e-3f;
    const float a = 1.0001f, b = 1e-4f * (float)iters;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v = fmaf(v, a, b);
    }
    acc = v;
  } else {  // FFMA2: one dependent chain
    uint64_t v, a, b;
    { const float af = 1.0001f, bf = 1e-4f * (float)iters, f = (float)threadIdx.x * 1e-3f;
      asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af)); asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf)); asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f)); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(a), "l"(b));
    }
    float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); acc = lo + hi;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 8 * 148); cudaMalloc(&sink, 4);
  const int iters = 2000;
  const char* names[4] = {"FFMA  x16 independent", "FFMA2 x16 independent", "FFMA  dependent chain", "FFMA2 dependent chain"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int warps : {1, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 1) k<1><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 2) k<2><<<1, warps * 32>>>(iters, cyc, sink);
        if (mode == 3) k<3><<<1, warps * 32>>>(iters, cyc, sink);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double instr = (double)iters * 16 * warps;
      printf("%s  warps=%2d  cycles=%8lld  warp-instr/clk/SM=%.3f  fp32 FMA lanes/clk/SM=%.1f  cycles/instr/warp=%.2f\n", names[mode], warps, c,
             instr / c, instr * 32 * ((mode & 1) ? 2 : 1) / c, (double)c / (iters * 16));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
