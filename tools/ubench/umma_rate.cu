// Micro-benchmark (development aid): cycles per tcgen05.mma (kind::f16, bf16 -> fp32, M=128, K=16) as a function of N, of the
// operand majors, and of where A lives (shared memory descriptor vs TMEM). Operands are whatever shared memory holds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int amn, int bmn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// mode 0: SS, A K-major, B K-major    1: SS, A MN-major, B MN-major   2: SS, A K-major, B MN-major   3: TS (A in TMEM), B MN-major
// The issue loop is fully unrolled over 16 MMAs with descriptors precomputed in registers, so that the issuing thread's own
// scalar instructions do not bound the measurement.
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) k(int n_outer, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = base, b_addr = base + 16 * 1024;
    constexpr uint32_t id = idesc(128, N, (MODE == 1) ? 1 : 0, (MODE == 0) ? 0 : 1);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ad[q] = (MODE == 1) ? desc_sw128(a_addr + q * 2048, 8192, 1024) : desc_sw128(a_addr + q * 32, 16, 1024);
      bd[q] = (MODE == 0) ? desc_sw128(b_addr + q * 32, 16, 1024) : desc_sw128(b_addr + q * 2048, 8192, 1024);
    }
    const long long t0 = clock64();
    for (int o = 0; o < n_outer; ++o) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t td = tm + (uint32_t)((i % NACC) * N);
        if (MODE < 3) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(td), "l"(ad[i & 3]), "l"(bd[i & 3]), "r"(id), "r"(1u) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                       ::"r"(td), "r"(tm + 256 + (uint32_t)(i & 3) * 8), "l"(bd[i & 3]), "r"(id), "r"(1u) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    cycles[0] = t1 - t0;
    cycles[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int MODE, int N, int NACC>
void run(long long* cyc, const char* name) {
  const int n_outer = 16;
  cudaFuncSetAttribute(k<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<MODE, N, NACC><<<1, 128, 64 * 1024>>>(n_outer, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s N=%d: error %s\n", name, N, cudaGetErrorString(e)); exit(1); }
  long long c[2];
  cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
  printf("%s nacc=%d M=128 N=%3d K=16: issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", name, NACC, N, (double)c[0] / (16 * n_outer),
         (double)c[1] / (16 * n_outer));
}

template <int MODE>
void run_mode(long long* cyc, const char* name) {
  run<MODE, 32, 1>(cyc, name); run<MODE, 32, 2>(cyc, name); run<MODE, 32, 4>(cyc, name);
  run<MODE, 64, 1>(cyc, name); run<MODE, 64, 2>(cyc, name);
  run<MODE, 128, 1>(cyc, name); run<MODE, 128, 2>(cyc, name);
  run<MODE, 192, 1>(cyc, name); run<MODE, 256, 1>(cyc, name);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 64);
  run_mode<0>(cyc, "SS A:K  B:K ");
  run_mode<1>(cyc, "SS A:MN B:MN");
  run_mode<2>(cyc, "SS A:K  B:MN");
  run_mode<3>(cyc, "TS A:tmem B:MN");
  return 0;
}
