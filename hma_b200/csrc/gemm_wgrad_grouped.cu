// The weight gradients of ONE ST block in one launch: dW_j[Mw_j, Nw_j] += G_j[tokens, Mw_j]^T . X_j[tokens, Nw_j] for up to
// eight Linears that share the token dimension (autograd transposes of attention.py:141,154; st_transformer.py:24-27;
// st_mask_git.py:70-75 — per layer: fc2, fc1, temporal proj / qkv, ModulateLayer.linear_out, spatial proj / qkv).
//
// Why grouped: launched one by one (gemm_wgrad.cu) each of these is a single wave of CTAs whose life is mostly prologue,
// pipeline fill and the red.global of a partial tile that only 1/37th .. 1/148th of the tokens went into — seven launches
// per layer at ~45 % of the tensor rate the shapes allow, 15 % of the training step. Here 148 persistent CTAs walk a list of
// (token slice, 128 x 256 tile) work items of ALL the group's matrices:
//   * item = slice * n_tiles + tile, so at any moment the CTAs work on the same few token slices and an operand slice read
//     by several tiles (every m-block of fc1 reads the same rows of its input) is served from L2;
//   * ~3 items per CTA with ~50 k-blocks each: the accumulator is double buffered in TMEM (2 x 256 columns), the
//     red.global epilogue of item i overlaps the contraction of item i+1, and a tile receives 13 partials instead of 37-148;
//   * every tcgen05.mma is M128 x N256 x K16, the shape that runs at the tensor pipe's full rate (measured 127 cycles).
// Operands are read "MN-major" exactly as in gemm_wgrad.cu: TMA lays [64 tokens x 64 channels] boxes down with the
// 128-byte swizzle and the descriptors walk them along the token axis.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kGrpMax = 8;
constexpr int kGBox = 64 * 64 * 2;            // one [64 tok x 64 ch] bf16 box = 8 KB
constexpr int kGBN = 256;                     // tile width
constexpr int kGStages = 4;
constexpr int kGGStage = 2 * kGBox;           // G: 128 channels
constexpr int kGXStage = (kGBN / 64) * kGBox; // X: 256 channels
constexpr int kGStageBytes = kGGStage + kGXStage;  // 48 KB

struct GrpMaps {
  CUtensorMap g[kGrpMax];
  CUtensorMap x[kGrpMax];
};

struct GrpParams {
  int count;
  int tokens;
  int chunk;      // tokens per slice, multiple of 64
  int n_slices;
  int n_tiles;    // over all problems
  int tile_start[kGrpMax + 1];
  int m_tiles[kGrpMax];
  float* dW[kGrpMax];
  long long ldw[kGrpMax];
};

struct GrpItem {
  int prob, m_blk, n_blk, tok0, kb;
};

__device__ __forceinline__ GrpItem grp_item(const GrpParams& p, int item) {
  GrpItem it;
  const int slice = item / p.n_tiles;
  const int tile = item - slice * p.n_tiles;
  int j = 0;
#pragma unroll
  for (int q = 1; q < kGrpMax; ++q)
    if (q < p.count && tile >= p.tile_start[q]) j = q;
  const int t = tile - p.tile_start[j];
  it.prob = j;
  it.m_blk = t % p.m_tiles[j];
  it.n_blk = t / p.m_tiles[j];
  it.tok0 = slice * p.chunk;
  int tok1 = it.tok0 + p.chunk;
  if (tok1 > p.tokens) tok1 = p.tokens;
  it.kb = (tok1 - it.tok0 + 63) / 64;
  return it;
}

__global__ void __launch_bounds__(256, 1)
gemm_wgrad_grouped_kernel(const __grid_constant__ GrpMaps maps, const __grid_constant__ GrpParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kGStages];
  __shared__ __align__(8) uint64_t bar_empty[kGStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = p.n_tiles * p.n_slices;

  if (warp == 0 && lane == 0) {
    for (int j = 0; j < p.count; ++j) {
      tma_prefetch_desc(&maps.g[j]);
      tma_prefetch_desc(&maps.x[j]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), 2 * kGBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const GrpItem it = grp_item(p, item);
        const CUtensorMap* mg = &maps.g[it.prob];
        const CUtensorMap* mx = &maps.x[it.prob];
        for (int kb = 0; kb < it.kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, (uint32_t)kGStageBytes);
          const uint32_t base = smem_base + stage * kGStageBytes;
          const int tok = it.tok0 + kb * 64;
#pragma unroll
          for (int i = 0; i < 2; ++i) tma_load_2d(base + i * kGBox, mg, full, it.m_blk * 128 + i * 64, tok);
#pragma unroll
          for (int j = 0; j < kGBN / 64; ++j) tma_load_2d(base + kGGStage + j * kGBox, mx, full, it.n_blk * kGBN + j * 64, tok);
          if (++stage == kGStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- UMMA issuer
    if (elect_one()) {
      constexpr uint32_t kIdesc = umma_idesc_bf16(128, kGBN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const GrpItem it = grp_item(p, item);
        const int as = n & 1;
        mbar_wait(smem_u32(&bar_tempty[as]), (uint32_t)((n >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kGBN);
        for (int kb = 0; kb < it.kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t g_addr = smem_base + stage * kGStageBytes;
          const uint32_t x_addr = g_addr + kGGStage;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(d_tmem, umma_desc_mnmajor(g_addr + k * 2048, kGBox), umma_desc_mnmajor(x_addr + k * 2048, kGBox), kIdesc,
                    (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == kGStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_tfull[as]));
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: partial tile -> red.global into dW
    const int ew = warp - 4;
    int n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const GrpItem it = grp_item(p, item);
      const int as = n & 1;
      mbar_wait(smem_u32(&bar_tfull[as]), (uint32_t)((n >> 1) & 1));
      tc_fence_after();
      const int m = it.m_blk * 128 + ew * 32 + lane;
      float* drow = p.dW[it.prob] + (size_t)m * p.ldw[it.prob] + (size_t)it.n_blk * kGBN;
#pragma unroll 1
      for (int c = 0; c < kGBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem_addr(tmem_base, (uint32_t)(ew * 32), (uint32_t)(as * kGBN + c * 32)), r);
        tmem_ld_wait();
        if (c == kGBN / 32 - 1) {  // the accumulator is in registers: hand it back before the atomics are issued
          tc_fence_before();
          mbar_arrive(smem_u32(&bar_tempty[as]));
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c * 32 + j),
                       "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                       "f"(__uint_as_float(r[j + 3]))
                       : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kGBN);
  }
}

}  // namespace hma

extern "C" int hma_gemm_wgrad_grouped(int count, const void* const* G, const long long* ldg, const void* const* X,
                                      const long long* ldx, int tokens, const int* Mw, const int* Nw, float* const* dW,
                                      const long long* ldw, void* stream_) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (count == 0 || tokens == 0) return 0;
  HMA_REQUIRE(count >= 1 && count <= kGrpMax, "gemm_wgrad_grouped: count=%d must be in [1,%d]", count, kGrpMax);
  HMA_REQUIRE(tokens > 0, "gemm_wgrad_grouped: bad token count %d", tokens);
  GrpMaps maps;
  GrpParams p;
  p.count = count;
  p.tokens = tokens;
  int tiles = 0;
  for (int j = 0; j < count; ++j) {
    HMA_REQUIRE(Mw[j] > 0 && Mw[j] % 128 == 0, "gemm_wgrad_grouped: Mw[%d]=%d must be a multiple of 128", j, Mw[j]);
    HMA_REQUIRE(Nw[j] > 0 && Nw[j] % kGBN == 0, "gemm_wgrad_grouped: Nw[%d]=%d must be a multiple of %d", j, Nw[j], kGBN);
    HMA_REQUIRE((ldw[j] % 4) == 0 && (reinterpret_cast<uintptr_t>(dW[j]) & 15) == 0, "gemm_wgrad_grouped: dW[%d] must be 16-byte aligned", j);
    int rc = hma_host::make_tmap_bf16_2d(&maps.g[j], G[j], (uint64_t)Mw[j], (uint64_t)tokens, (uint64_t)ldg[j] * 2, 64, 64);
    if (rc) return rc;
    rc = hma_host::make_tmap_bf16_2d(&maps.x[j], X[j], (uint64_t)Nw[j], (uint64_t)tokens, (uint64_t)ldx[j] * 2, 64, 64);
    if (rc) return rc;
    p.tile_start[j] = tiles;
    p.m_tiles[j] = Mw[j] / 128;
    tiles += (Mw[j] / 128) * (Nw[j] / kGBN);
    p.dW[j] = dW[j];
    p.ldw[j] = ldw[j];
  }
  for (int j = count; j < kGrpMax; ++j) {
    maps.g[j] = maps.g[0];
    maps.x[j] = maps.x[0];
    p.tile_start[j] = tiles;
    p.m_tiles[j] = 1;
    p.dW[j] = nullptr;
    p.ldw[j] = 0;
  }
  p.tile_start[count] = tiles;
  p.n_tiles = tiles;
  // Slices: items = tiles x slices should fill a whole number of waves of the persistent CTAs (a little under, never a
  // little over), with slices of at least 512 tokens so that an item is much longer than its pipeline fill.
  const int sms = hma_host::sm_count();
  const int max_slices = tokens / 512 > 0 ? (tokens / 512 < 64 ? tokens / 512 : 64) : 1;
  auto eff_of = [&](int s) {
    const int items = tiles * s;
    const int waves = (items + sms - 1) / sms;
    return (double)items / ((double)waves * sms);
  };
  double best_eff = 0.0;
  for (int s = 1; s <= max_slices; ++s) best_eff = eff_of(s) > best_eff ? eff_of(s) : best_eff;
  int best_slices = 0;
  for (int s = 1; s <= max_slices && best_slices == 0; ++s)  // fewest partial tiles among the well-filled choices with
    if (eff_of(s) >= best_eff - 0.015 && tiles * s >= 2 * sms) best_slices = s;  // >= 2 items per CTA (epilogue overlap)
  for (int s = 1; s <= max_slices && best_slices == 0; ++s)
    if (eff_of(s) >= best_eff - 0.015) best_slices = s;
  int chunk = (tokens + best_slices - 1) / best_slices;
  chunk = (chunk + 63) / 64 * 64;
  p.chunk = chunk;
  p.n_slices = (tokens + chunk - 1) / chunk;
  const int n_items = p.n_tiles * p.n_slices;
  const int grid = n_items < sms ? n_items : sms;
  constexpr size_t smem = 1024 + (size_t)kGStages * kGStageBytes;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(gemm_wgrad_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  HMA_CHECK_CUDA(hma_host::launch_pdl(gemm_wgrad_grouped_kernel, dim3(grid), dim3(256), smem, stream, maps, p));
  return 0;
}
