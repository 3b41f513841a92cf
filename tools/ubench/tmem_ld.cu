// Micro-benchmark (development aid): throughput of tcgen05.ld 32x32b.x32 and of MUFU ex2 per SM, as a function of the
// number of warps. Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld tools/ubench/tmem_ld.cu && /tmp/tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {  // TMEM loads only
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      const uint32_t addr = base + (uint32_t)((i * 32) & 511) ;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(addr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[i & 31] & 0x3fffffffu);
    }
  } else if (MODE == 1) {  // MUFU ex2: 32 per iteration, independent
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x + j) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += v[j];
  } else {  // FFMA: 32 independent chains per iteration
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x + j) * 1e-3f;
    const float a = 1.0001f, b = 1e-4f * (float)iters;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += v[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 8 * 256);
  cudaMalloc(&sink, 4 * 1024 * 256);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {1, 2, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(iters, cyc, sink);
        else if (mode == 1) k<1><<<1, warps * 32>>>(iters, cyc, sink);
        else k<2><<<1, warps * 32>>>(iters, cyc, sink);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_warp_instr = (double)c / iters;
      if (mode == 0)
        printf("tcgen05.ld x32: warps=%2d  %7.1f cyc per ld per warp -> %7.1f B/cyc/SM\n", warps, per_warp_instr,
               warps * 4096.0 / per_warp_instr);
      else
        printf("%s: warps=%2d  %7.2f cyc per warp-instruction -> %6.1f lanes/cyc/SM\n", mode == 1 ? "ex2 " : "ffma", warps,
               per_warp_instr / 32.0, warps * 32.0 * 32.0 / per_warp_instr);
    }
  }
  return 0;
}
