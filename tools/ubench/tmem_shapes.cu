// Micro-benchmark (development aid): tcgen05.ld throughput per SM for the different load shapes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define R32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define O32 "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), \
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), \
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])

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
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    const uint32_t addr = base + (uint32_t)((i * 32) & 255);
    if (MODE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " R32 ", [%32];" : O32 : "r"(addr) : "memory");
    if (MODE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " R32 ", [%32];" : O32 : "r"(addr) : "memory");
    if (MODE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " R32 ", [%32];" : O32 : "r"(addr) : "memory");
    if (MODE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " R32 ", [%32];" : O32 : "r"(addr) : "memory");
    if (MODE == 4) {  // two x32 loads in flight before the wait
      uint32_t q[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " R32 ", [%32];" : O32 : "r"(addr) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                   : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
                     "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]),
                     "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]),
                     "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                   : "r"(addr + 32) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(q[i & 31] & 0x3fffffffu);
    }
    if (MODE == 5) {  // 32x32b.x8: 1 KB per instruction
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(r[i & 7] & 0x3fffffffu);
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
  const char* names[6] = {"32x32b.x32", "16x256b.x8", "16x128b.x16", "16x64b.x32", "2x 32x32b.x32 in flight", "32x32b.x8"};
  const double bytes[6] = {4096, 4096, 4096, 4096, 8192, 1024};
  for (int mode = 0; mode < 6; ++mode) {
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: k<0><<<1, warps * 32>>>(iters, cyc, sink); break;
          case 1: k<1><<<1, warps * 32>>>(iters, cyc, sink); break;
          case 2: k<2><<<1, warps * 32>>>(iters, cyc, sink); break;
          case 3: k<3><<<1, warps * 32>>>(iters, cyc, sink); break;
          case 4: k<4><<<1, warps * 32>>>(iters, cyc, sink); break;
          case 5: k<5><<<1, warps * 32>>>(iters, cyc, sink); break;
        }
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: error %s\n", names[mode], cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-26s warps=%2d  %7.1f cyc/iter/warp -> %7.1f B/cyc/SM\n", names[mode], warps, (double)c / iters,
             warps * bytes[mode] * iters / (double)c);
    }
  }
  return 0;
}
