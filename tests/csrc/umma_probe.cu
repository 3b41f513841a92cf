// Descriptor probe: a single-CTA kernel that TMA-loads two bf16 matrices into shared memory and
// issues tcgen05.mma with descriptor fields chosen by the host, then dumps the fp32 accumulator.
// Used by tests/test_umma_probe_gpu.py to pin down (on the real B200) every shared-memory
// descriptor variant the attention kernels rely on before those kernels depend on it.
// TEST INFRASTRUCTURE: built into tests/libhma_b200_probe.so (hma_b200.build.build_probe), not part of the product ABI.
#include "../../hma_b200/csrc/common.cuh"

extern "C" int hma_umma_probe(const void* A, long long lda, const void* B, long long ldb, const int* params, float* out,
                              void* stream);

namespace hma {

struct ProbeParams {
  int a_rows, a_boxes, b_rows, b_boxes;  // boxes of [rows x box_inner] laid down back to back
  int box_inner;                         // 64 (SW128) or 32 (SW64)
  int layout_type;                       // UMMA layout type field: 2 = SW128, 4 = SW64
  int a_major, b_major;                  // 0 = K, 1 = MN
  int a_off, b_off;                      // byte offsets added to the operand start address
  int a_lbo, a_sbo, b_lbo, b_sbo;        // descriptor byte offsets
  int ksteps, a_kstep, b_kstep;          // number of k16 MMAs and byte advance per step
  int N;
  float* out;                            // [128, N]
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo, int layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(((uint32_t)lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>(((uint32_t)sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int row_bytes = p.box_inner * 2;
  const uint32_t a_bytes = (uint32_t)(p.a_rows * row_bytes * p.a_boxes);
  const uint32_t b_bytes = (uint32_t)(p.b_rows * row_bytes * p.b_boxes);
  const uint32_t smemA = smem_base;
  const uint32_t smemB = (smem_base + a_bytes + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar_full), a_bytes + b_bytes);
    for (int i = 0; i < p.a_boxes; ++i)
      tma_load_2d(smemA + i * p.a_rows * row_bytes, &tmA, smem_u32(&bar_full), i * p.box_inner, 0);
    for (int i = 0; i < p.b_boxes; ++i)
      tma_load_2d(smemB + i * p.b_rows * row_bytes, &tmB, smem_u32(&bar_full), i * p.box_inner, 0);
    mbar_wait(smem_u32(&bar_full), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.N, p.a_major, p.b_major);
    for (int k = 0; k < p.ksteps; ++k) {
      umma_ss(tmem_base, make_desc(smemA + p.a_off + k * p.a_kstep, p.a_lbo, p.a_sbo, p.layout_type),
              make_desc(smemB + p.b_off + k * p.b_kstep, p.b_lbo, p.b_sbo, p.layout_type), idesc, (uint32_t)(k != 0));
    }
    umma_commit(smem_u32(&bar_done));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_done), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < p.N; c += 16) {
    uint32_t r[16];
    tmem_ld_x16(tmem_addr(tmem_base, (uint32_t)(warp * 32), (uint32_t)c), r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j)
      if (c + j < p.N) p.out[(size_t)row * p.N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace hma

// params: int[18] = {a_rows,a_boxes,b_rows,b_boxes,box_inner,layout_type,a_major,b_major,a_off,b_off,
//                    a_lbo,a_sbo,b_lbo,b_sbo,ksteps,a_kstep,b_kstep,N}
extern "C" int hma_umma_probe(const void* A, long long lda, const void* B, long long ldb, const int* params,
                              float* out, void* stream_) {
  using namespace hma;
  ProbeParams p;
  p.a_rows = params[0]; p.a_boxes = params[1]; p.b_rows = params[2]; p.b_boxes = params[3];
  p.box_inner = params[4]; p.layout_type = params[5]; p.a_major = params[6]; p.b_major = params[7];
  p.a_off = params[8]; p.b_off = params[9]; p.a_lbo = params[10]; p.a_sbo = params[11];
  p.b_lbo = params[12]; p.b_sbo = params[13]; p.ksteps = params[14]; p.a_kstep = params[15];
  p.b_kstep = params[16]; p.N = params[17];
  p.out = out;
  HMA_REQUIRE(p.box_inner == 64 || p.box_inner == 32, "probe: box_inner must be 64 or 32");
  HMA_REQUIRE(p.N >= 16 && p.N <= 256 && p.N % 16 == 0, "probe: bad N");
  const int sw = p.box_inner * 2;
  CUtensorMap tmA, tmB;
  int rc = hma_host::make_tmap_bf16_2d_sw(&tmA, A, (uint64_t)p.a_boxes * p.box_inner, (uint64_t)p.a_rows,
                                          (uint64_t)lda * 2, p.box_inner, p.a_rows, sw);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d_sw(&tmB, B, (uint64_t)p.b_boxes * p.box_inner, (uint64_t)p.b_rows,
                                      (uint64_t)ldb * 2, p.box_inner, p.b_rows, sw);
  if (rc) return rc;
  const size_t smem = 3072 + (size_t)(p.a_rows * p.a_boxes + p.b_rows * p.b_boxes) * sw;
  HMA_REQUIRE(smem <= 200 * 1024, "probe: tiles too large");
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream_)>>>(tmA, tmB, p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
