// dW[Mw,Nw] += G[tokens,Mw]^T . X[tokens,Nw]: the weight-gradient contraction of every Linear on
// the path (autograd transposes of attention.py:141,154; st_transformer.py:24-27;
// st_mask_git.py:70-75,681-683). The reduction runs over tokens, so both operands are read
// "MN-major": TMA lays [64 tokens x 64 channels] boxes down with the 128-byte swizzle and the
// UMMA descriptors walk them along the token axis.
//
// One CTA = one 128 x BNW tile of dW over one slice of the tokens (split-K); partial tiles are
// added into fp32 dW with vector red.global. Same warp roles as gemm_nt.cu.
#include <cstdlib>
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct WgradParams {
  int tokens, Mw, Nw;
  int chunk;  // tokens per split, multiple of 64
  float* dW;
  long long ldw;
};

constexpr int kBox = 64 * 64 * 2;  // one [64 tok x 64 ch] bf16 box = 8 KB
// ~192 KB of operand stages per CTA whatever the tile width
template <int BNW> struct WStages { static constexpr int value = (192 * 1024) / ((2 + BNW / 64) * kBox); };

template <int BNW>
__global__ void __launch_bounds__(256, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  constexpr int kWStages = WStages<BNW>::value;
  constexpr int kGStage = 2 * kBox;
  constexpr int kXStage = (BNW / 64) * kBox;
  constexpr int kStage = kGStage + kXStage;
  constexpr uint32_t kIdesc = umma_idesc_bf16(128, BNW, 1, 1);

  const int tok0 = blockIdx.y * p.chunk;
  int tok1 = tok0 + p.chunk;
  if (tok1 > p.tokens) tok1 = p.tokens;
  if (tok0 >= tok1) return;  // uniform for the whole CTA, before any barrier / TMEM state exists
  const int KB = (tok1 - tok0 + 63) / 64;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWStages];
  __shared__ __align__(8) uint64_t bar_empty[kWStages];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.Mw / 128;
  const int m_blk = blockIdx.x % m_tiles;
  const int n_blk = blockIdx.x / m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), BNW);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&bar_full[stage]);
        mbar_expect_tx(full, (uint32_t)kStage);
        const uint32_t base = smem_base + stage * kStage;
        const int tok = tok0 + kb * 64;
#pragma unroll
        for (int i = 0; i < 2; ++i) tma_load_2d(base + i * kBox, &tmG, full, m_blk * 128 + i * 64, tok);
#pragma unroll
        for (int j = 0; j < BNW / 64; ++j)
          tma_load_2d(base + kGStage + j * kBox, &tmX, full, n_blk * BNW + j * 64, tok);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t g_addr = smem_base + stage * kStage;
        const uint32_t x_addr = g_addr + kGStage;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_ss(tmem_base, umma_desc_mnmajor(g_addr + k * 2048, kBox), umma_desc_mnmajor(x_addr + k * 2048, kBox),
                  kIdesc, (uint32_t)((kb | k) != 0));
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(&bar_done));
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int m = m_blk * 128 + ew * 32 + lane;
    float* drow = p.dW + (size_t)m * p.ldw + (size_t)n_blk * BNW;
#pragma unroll 1
    for (int c = 0; c < BNW / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_addr(tmem_base, (uint32_t)(ew * 32), (uint32_t)(c * 32)), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c * 32 + j),
                     "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                     "f"(__uint_as_float(r[j + 3]))
                     : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BNW);
  }
}

template <int BNW>
static int launch_wgrad(const void* G, long long ldg, const void* X, long long ldx, WgradParams p,
                        cudaStream_t stream) {
  CUtensorMap tmG, tmX;
  int rc = hma_host::make_tmap_bf16_2d(&tmG, G, (uint64_t)p.Mw, (uint64_t)p.tokens, (uint64_t)ldg * 2, 64, 64);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d(&tmX, X, (uint64_t)p.Nw, (uint64_t)p.tokens, (uint64_t)ldx * 2, 64, 64);
  if (rc) return rc;
  constexpr size_t smem = 1024 + (size_t)WStages<BNW>::value * (2 * kBox + (BNW / 64) * kBox);
  auto kern = gemm_wgrad_kernel<BNW>;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int tiles = (p.Mw / 128) * (p.Nw / BNW);
  int splits = hma_host::sm_count() / tiles;
  if (splits < 1) splits = 1;
  int chunk = (p.tokens + splits - 1) / splits;
  chunk = (chunk + 63) / 64 * 64;
  splits = (p.tokens + chunk - 1) / chunk;
  p.chunk = chunk;
  HMA_CHECK_CUDA(hma_host::launch_pdl(kern, dim3(tiles, splits), dim3(256), smem, stream, tmG, tmX, p));
  return 0;
}

}  // namespace hma

extern "C" int hma_gemm_wgrad(const void* G, long long ldg, const void* X, long long ldx, int tokens, int Mw,
                              int Nw, float* dW, long long ldw, void* stream_) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (tokens == 0) return 0;
  HMA_REQUIRE(tokens > 0 && Mw > 0 && Nw > 0, "gemm_wgrad: bad shape tokens=%d Mw=%d Nw=%d", tokens, Mw, Nw);
  HMA_REQUIRE(Mw % 128 == 0, "gemm_wgrad: Mw=%d must be a multiple of 128", Mw);
  HMA_REQUIRE(Nw % 64 == 0, "gemm_wgrad: Nw=%d must be a multiple of 64", Nw);
  HMA_REQUIRE((ldw % 4) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0, "gemm_wgrad: dW must be 16-byte aligned");
  WgradParams p;
  p.tokens = tokens; p.Mw = Mw; p.Nw = Nw; p.chunk = 0; p.dW = dW; p.ldw = ldw;
  // Tile width: every CTA adds its 128 x BNW partial tile into dW with L2 atomics (measured ~0.76 T float adds/s,
  // the largest single cost of the 128 x 256 version), while a narrower tile re-reads the operands through L2
  // more often. Pick the width that minimises max(HBM time, L2 time) + atomic time.
  const int sms = hma_host::sm_count();
  int best = 0;
  double best_t = 0.0;
  for (int bnw : {64, 128, 256}) {
    if (Nw % bnw != 0) continue;
    const int tiles = (Mw / 128) * (Nw / bnw);
    int splits = sms / tiles;
    if (splits < 1) splits = 1;
    const double ctas = (double)tiles * splits;
    const double t_atom = ctas * 128.0 * bnw / 0.76e12;
    const double t_l2 = 2.0 * tokens * ((double)Mw * (Nw / bnw) + (double)Nw * (Mw / 128)) / 20e12;
    const double t_hbm = 2.0 * tokens * ((double)Mw + Nw) / 6.5e12;
    const double t = (t_l2 > t_hbm ? t_l2 : t_hbm) + t_atom;
    if (best == 0 || t < best_t) { best = bnw; best_t = t; }
  }
  if (best == 256) return launch_wgrad<256>(G, ldg, X, ldx, p, stream);
  if (best == 128) return launch_wgrad<128>(G, ldg, X, ldx, p, stream);
  return launch_wgrad<64>(G, ldg, X, ldx, p, stream);
}
