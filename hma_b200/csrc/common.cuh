// Shared device/host primitives for the sm_100a kernels: mbarrier, TMA, tcgen05 (UMMA + TMEM),
// shared-memory descriptor construction and the software side of the 128-byte swizzle.
// Everything here is inline PTX; nothing is borrowed from a library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hma {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ------------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100: fma / mul / add .f32x2 -> FFMA2 / FMUL2 / FADD2). One warp instruction does 64 fp32 operations at
// the latency (4 cycles) and issue cost of a scalar one (measured, tools/ubench/ffma2.cu: the pipe's peak lanes/clk is the
// same, so this does nothing for throughput-bound code) — the epilogues and softmax stages here run a handful of warps and are
// bound by issue slots and dependent latency, where halving the instruction count is what counts. nvcc never emits them
// from scalar code. A pair lives in a 64-bit register: {lo, hi}.
// ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 dup2(float x) { return pk2(x, x); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16(lo, hi);
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// Bounded wait: a protocol bug traps (surfacing as a CUDA error on the host) instead of hanging the GPU box. The bound
// is WALL TIME (%globaltimer, checked every 4096 failed polls), not a spin count: a kernel that is merely slow — under a
// profiler's replay passes, a debugger, time-slicing or preemption — must not trap.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kMbarTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0xfffu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kMbarTimeoutNs) __trap();
    }
  }
}
// ------------------------------------------------------------------------------------------
// Thread-block clusters: distributed shared memory between the CTAs of a cluster.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster (release / acquire at cluster scope)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the shared::cluster address of this CTA's shared-memory address `addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// Asynchronous store of a float pair into the shared memory of a CTA of the cluster; completion is counted in BYTES on an
// mbarrier of that same CTA (mbarrier.expect_tx there), so neither side needs a release fence — a release at cluster scope
// would first drain every global store the thread has in flight (measured: ~2700 cycles per tile in the LN epilogue).
__device__ __forceinline__ void st_async_f32x2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "f"(a),
               "f"(b), "r"(remote_bar)
               : "memory");
}
// arrive on an mbarrier of another CTA of the cluster; releases this thread's earlier (remote) writes at cluster scope
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// wait on a local mbarrier that CTAs of the cluster arrive on (acquire at cluster scope); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0xfffu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kMbarTimeoutNs) __trap();
    }
  }
}
// generic-proxy smem writes -> visible to the async proxy (TMA / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch: kernels are launched with programmatic stream serialization
// (hma_host::launch_pdl), so a kernel's CTAs may start while the previous kernel in the stream is
// still draining. Everything before pdl_wait() (barrier init, TMEM allocation, descriptor
// prefetch, shared-memory zeroing) overlaps the predecessor's tail; pdl_wait() returns once the
// predecessor grid has completed and its memory is visible. NO global memory may be read or
// written before it. pdl_launch_dependents() lets the next kernel's CTAs be scheduled early.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// Development aid (build with HMA_B200_TIMELINE=1): CTA 0 stamps clock64() into a per-source-file
// table, event `ev` (< 16), slot `i` (< 64); a file that uses it also emits a reader entry point with
// HMA_DEFINE_TIMELINE_READER(hma_timeline_<file>) (see tools/timeline.py). Compiled out otherwise.
// ------------------------------------------------------------------------------------------
#ifdef HMA_TIMELINE
static __device__ long long g_timeline[16 * 64];
#define HMA_DEFINE_TIMELINE_READER(fn)                                                              \
  extern "C" int fn(long long* host_dst) {                                                          \
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;                                          \
    return cudaMemcpyFromSymbol(host_dst, hma::g_timeline, sizeof(long long) * 16 * 64) == cudaSuccess ? 0 : -1; \
  }
#define HMA_TL(ev, i)                                                                   \
  do {                                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (i) < 64) g_timeline[(ev) * 64 + (i)] = clock64(); \
  } while (0)
#else
#define HMA_TL(ev, i) do { } while (0)
#define HMA_DEFINE_TIMELINE_READER(fn)
#endif

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiles
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// TMEM allocation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// UMMA descriptors (layouts follow the canonical forms documented for sm_100 matrix descriptors)
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle.
//   bits [0,14)  start address >> 4
//   bits [16,30) leading-dimension byte offset >> 4
//   bits [32,46) stride-dimension byte offset >> 4
//   bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// K-major operand tile: rows of 128 bytes (64 bf16 along K), 8-row groups 1024 bytes apart.
// One descriptor covers UMMA_K = 16 elements (32 bytes); advance K by adding 32 bytes to saddr.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
  return umma_desc_sw128(saddr, 16, 1024);
}
// MN-major operand tile as TMA lays it down for a [k rows x 64 ch] box: each k row is 128 bytes
// (64 bf16 along M/N), 8-row k groups are 1024 bytes apart (SBO), the next 64 channels start
// `mn_atom_stride` bytes later (LBO). Advance K (16 rows) by adding 2048 bytes to saddr.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t saddr, uint32_t mn_atom_stride) {
  return umma_desc_sw128(saddr, mn_atom_stride, 1024);
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4)                                   // D format: f32
         | (1u << 7)                                 // A format: bf16
         | (1u << 10)                                // B format: bf16
         | (static_cast<uint32_t>(a_mn_major) << 15) // A major
         | (static_cast<uint32_t>(b_mn_major) << 16) // B major
         | (static_cast<uint32_t>(N >> 3) << 17)     // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);    // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued UMMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// TMEM -> registers. 32x32b shape: lane i of the warp reads TMEM lane (base_lane + i),
// N consecutive 32-bit columns. The warp may only touch lanes [32*(warp%4), 32*(warp%4)+32).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// TMEM address = (lane << 16) | column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// ------------------------------------------------------------------------------------------
// Software view of the 128-byte swizzle, for tiles written by ordinary st.shared
// (K-major: `row` indexes M/N, `col` indexes K inside one 64-element panel;
//  MN-major: `row` indexes K, `col` indexes M/N inside one 64-element panel).
// The tile base must be 1024-byte aligned.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col_elem) {
  const uint32_t chunk = (col_elem >> 3) ^ (row & 7u);
  return (row >> 3) * 1024u + (row & 7u) * 128u + chunk * 16u + (col_elem & 7u) * 2u;
}

// ------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float z) {
  return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
}
__device__ __forceinline__ float dgelu_erf(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
  return cdf + z * pdf;
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu(float z) { return z * fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * z)); }
__device__ __forceinline__ float dsilu(float z) {
  const float s = fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * z));
  return s * (1.0f + z * (1.0f - s));
}

}  // namespace hma

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace hma_host {

void set_error(const char* fmt, ...);
const char* last_error();

// 2-D bf16 tensor map with 128-byte swizzle; `inner` is the contiguous dimension.
// box_inner must be 64 (= 128 bytes); out-of-bounds elements read as zero.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

int make_tmap_bf16_2d_sw(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

int make_tmap_bf16_3d_sw64(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b1, uint32_t b2);

int sm_count();
int current_device();
// One flag per CUDA device: cudaFuncSetAttribute and friends are per-device state, and a process may drive several
// devices (idempotent values; racing threads set the same flag).
struct PerDeviceFlag {
  bool done[64] = {};
  bool& get() { return done[current_device() & 63]; }
};
bool pdl_enabled();  // HMA_B200_NO_PDL=1 turns programmatic dependent launch off (A/B measurements)

// Launch `kern` with programmatic stream serialization; the kernel MUST call hma::pdl_wait() before it
// touches global memory.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// As launch_pdl, for kernels whose CTAs form thread-block clusters of `cluster_x` consecutive blocks (gridDim.x % cluster_x == 0).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster_x,
                                      cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define HMA_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      hma_host::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                          cudaGetErrorString(_e));                                        \
      return -100 - static_cast<int>(_e);                                                 \
    }                                                                                     \
  } while (0)

#define HMA_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      hma_host::set_error(__VA_ARGS__);                                                   \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

}  // namespace hma_host
