// Bidirectional attention over the n (<= 320) tokens of one frame, head_dim 32:
//   O = softmax(scale * Q K^T) V      (reference: attention.py:37-61 / 139-155 with causal=False,
//                                      called from st_transformer.py:85-86)
// q, k, v are column slices of the bf16 [tokens, 3*d] output of the QKV projection.
//
// sm_100a design (one CTA per (frame, head), two CTAs resident per SM):
//   * TMA loads Q, K, V of the frame/head as 64-byte-swizzled [rows x 32] tiles (60 KB).
//   * keys are processed in two blocks (<= 192 and <= 128 keys): S = Q K^T goes to TMEM with one
//     tcgen05.mma pair per block; four softmax warps (one query row per thread, tcgen05.ld)
//     compute max / exp2 / sum and write P as bf16 into 128-byte-swizzled shared memory;
//     P V accumulates in TMEM (separate accumulator per key block, merged in the epilogue with
//     the usual log-sum-exp weights, so no TMEM rescale pass is needed).
//   * 256 TMEM columns and 109 KB shared memory per CTA: the second resident CTA's MMAs overlap
//     this CTA's softmax.
// Also writes the per-row log2-sum-exp needed by the backward kernel.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct AttnFwdParams {
  int n;          // tokens per frame
  int na, nb;     // key block sizes (na + nb == n)
  int box_rows;   // TMA box rows (divides n)
  int q_col, k_col, v_col;  // column of head 0 inside the qkv matrix
  float scale_log2;         // attn scale * log2(e)
  __nv_bfloat16* out;       // [tokens, ldo], head h at column h*32
  long long ldo;
  float* lse;               // [frames, heads, n] (log2 domain) or null
  int heads;
};

constexpr int kHd = 32;
constexpr int kRowB = 64;                   // bytes per q/k/v row (32 bf16)
constexpr int kMaxN = 320;
constexpr int kQKVBytes = kMaxN * kRowB;    // 20 KB per operand
constexpr int kPPanel = 128 * 128;          // one [128 x 64] bf16 panel, 16 KB
constexpr int kPBytes = 3 * kPPanel;

__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;  // SWIZZLE_64B
  return d;
}

// One key block of one query row is shared by TWO threads (warps w and w + 4 own the same TMEM lane quarter): each
// takes half of the block's 32-column chunks, [c_lo, c_hi). Pass 1: maximum over the own columns.
template <bool kFull>
__device__ __forceinline__ float block_max_t(uint32_t tmem_row_s, int c_lo, int c_hi, int ncols) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(tmem_row_s + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (kFull || c + j < ncols) {  // ncols is a multiple of 16: groups of four never straddle
        m0 = fmaxf(m0, __uint_as_float(r[j]));
        m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
        m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
      }
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// Pass 2: P = exp2(scale * s - mb) of the own columns as bf16 into the 128-byte-swizzled tile; returns their sum.
template <bool kFull>
__device__ __forceinline__ float block_exp_t(uint32_t tmem_row_s, int c_lo, int c_hi, int ncols, float scale_log2, float mb,
                                             uint32_t p_smem, int row) {
  float l0 = 0.f, l1 = 0.f;
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(tmem_row_s + c, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float p0 = fast_ex2(fmaf(__uint_as_float(r[j]), scale_log2, -mb));
      float p1 = fast_ex2(fmaf(__uint_as_float(r[j + 1]), scale_log2, -mb));
      if (!kFull && c + j >= ncols) { p0 = 0.f; p1 = 0.f; }
      l0 += p0;
      l1 += p1;
      pk[j >> 1] = pack_bf16(p0, p1);
    }
    const uint32_t panel = p_smem + (uint32_t)(c >> 6) * kPPanel;
    const int col0 = c & 63;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t addr = panel + sw128_offset((uint32_t)row, (uint32_t)(col0 + q * 8));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                   "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                   : "memory");
    }
  }
  return l0 + l1;
}
// barrier between the two warps that share a TMEM lane quarter (named barriers 1..4)
__device__ __forceinline__ void pair_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

constexpr int kSoftmaxWarps = 8;
constexpr int kFwdThreads = kSoftmaxWarps * 32 + 32;

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_spatial_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_s, bar_p, bar_o, bar_oread;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_xm[2][128], s_xl[2][128];  // per-row partial max / sum of the two column halves

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + kQKVBytes;
  const uint32_t sV = sK + kQKVBytes;
  const uint32_t sP = sV + kQKVBytes;  // 61440 = 60 * 1024: still 1024-aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int frame = blockIdx.x / p.heads;
  const int head = blockIdx.x % p.heads;
  const int n = p.n, na = p.na, nb = p.nb;
  const int ntiles = (n + 127) / 128;
  const int row0 = frame * n;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), kSoftmaxWarps * 32);
    mbar_init(smem_u32(&bar_o), 1);
    mbar_init(smem_u32(&bar_oread), 128);
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps) {
    tmem_alloc(smem_u32(&tmem_base_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // several waves of CTAs: dependents are released by CTA exit, not early
  const uint32_t tS = tmem_base;
  const uint32_t tOa = tmem_base + 192;
  const uint32_t tOb = tmem_base + 224;

  if (warp == kSoftmaxWarps) {
    if (elect_one()) {
      // ------------------------------------------------ loads
      const uint32_t bl = smem_u32(&bar_load);
      mbar_expect_tx(bl, (uint32_t)(3 * n * kRowB));
      for (int r = 0; r < n; r += p.box_rows) {
        tma_load_2d(sQ + r * kRowB, &tmQKV, bl, p.q_col + head * kHd, row0 + r);
        tma_load_2d(sK + r * kRowB, &tmQKV, bl, p.k_col + head * kHd, row0 + r);
        tma_load_2d(sV + r * kRowB, &tmQKV, bl, p.v_col + head * kHd, row0 + r);
      }
      mbar_wait(bl, 0);
      tc_fence_after();
      // ------------------------------------------------ MMA issue
      const uint32_t idesc_sa = umma_idesc_bf16(128, na, 0, 0);
      const uint32_t idesc_sb = umma_idesc_bf16(128, nb > 0 ? nb : 16, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, kHd, 0, 1);
      uint32_t pp = 0;
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t q_addr = sQ + (uint32_t)t * 128 * kRowB;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_ss(tS, desc_sw64(q_addr + k * 32, 16, 512), desc_sw64(sK + k * 32, 16, 512), idesc_sa, (uint32_t)k);
        umma_commit(smem_u32(&bar_s));
        mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
        tc_fence_after();
        if (t > 0) {
          mbar_wait(smem_u32(&bar_oread), (uint32_t)((t - 1) & 1));
          tc_fence_after();
        }
        for (int kk = 0; kk < na / 16; ++kk)
          umma_ss(tOa, umma_desc_kmajor(sP + (uint32_t)(kk >> 2) * kPPanel + (uint32_t)(kk & 3) * 32),
                  desc_sw64(sV + (uint32_t)kk * 16 * kRowB, 2048, 512), idesc_pv, (uint32_t)(kk != 0));
        if (nb > 0) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_ss(tS, desc_sw64(q_addr + k * 32, 16, 512), desc_sw64(sK + (uint32_t)na * kRowB + k * 32, 16, 512),
                    idesc_sb, (uint32_t)k);
          umma_commit(smem_u32(&bar_s));
          mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
          tc_fence_after();
          for (int kk = 0; kk < nb / 16; ++kk)
            umma_ss(tOb, umma_desc_kmajor(sP + (uint32_t)(kk >> 2) * kPPanel + (uint32_t)(kk & 3) * 32),
                    desc_sw64(sV + (uint32_t)(na + kk * 16) * kRowB, 2048, 512), idesc_pv, (uint32_t)(kk != 0));
        }
        umma_commit(smem_u32(&bar_o));
      }
    }
  } else {
    // -------------------------------------------------- softmax (warps 0-7) + epilogue (warps 0-3)
    // Warps w and w + 4 share the query rows of TMEM lane quarter w % 4 and split the key columns of every block:
    // with one row per thread and 320 keys the arithmetic of a single warp per quarter was latency-bound.
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;  // row inside the 128-query tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    auto softmax_half = [&](int ncols, float& m_out, float& l_out, bool live) {
      const int chunks = (ncols + 31) >> 5;
      const int c_lo = half == 0 ? 0 : ((chunks + 1) >> 1) * 32;
      const int c_hi = half == 0 ? ((chunks + 1) >> 1) * 32 : chunks * 32;
      const bool full = (ncols & 31) == 0;
      float m_part = -INFINITY;
      if (live) m_part = full ? block_max_t<true>(tS + lane_addr, c_lo, c_hi, ncols) : block_max_t<false>(tS + lane_addr, c_lo, c_hi, ncols);
      s_xm[half][row] = m_part;
      pair_sync(quarter);
      const float m = fmaxf(s_xm[0][row], s_xm[1][row]);
      float l_part = 0.f;
      if (live) {
        const float mb = m * p.scale_log2;
        l_part = full ? block_exp_t<true>(tS + lane_addr, c_lo, c_hi, ncols, p.scale_log2, mb, sP, row)
                      : block_exp_t<false>(tS + lane_addr, c_lo, c_hi, ncols, p.scale_log2, mb, sP, row);
      }
      s_xl[half][row] = l_part;
      pair_sync(quarter);
      m_out = m;
      l_out = s_xl[0][row] + s_xl[1][row];
    };
    uint32_t ps = 0, po = 0;
    for (int t = 0; t < ntiles; ++t) {
      float ma, la, mb = -INFINITY, lb = 0.f;
      // rows past the frame (second half of the last query tile) are never stored: their warps skip the
      // exp work and leave stale, finite-or-not P rows behind (a P row only feeds its own O row)
      const bool live = t * 128 + quarter * 32 < n;  // warp-uniform, same for both warps of a pair
      mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
      tc_fence_after();
      softmax_half(na, ma, la, live);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));
      if (nb > 0) {
        mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
        tc_fence_after();
        softmax_half(nb, mb, lb, live);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_p));
      }
      if (half != 0) continue;  // the epilogue of a row is done by its first thread
      mbar_wait(smem_u32(&bar_o), po); po ^= 1u;
      tc_fence_after();
      uint32_t oa[32], ob[32];
      tmem_ld_x32(tOa + lane_addr, oa);
      if (nb > 0) tmem_ld_x32(tOb + lane_addr, ob);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_oread));

      const float m = fmaxf(ma, mb);
      const float wa = fast_ex2((ma - m) * p.scale_log2);
      const float wb = nb > 0 ? fast_ex2((mb - m) * p.scale_log2) : 0.f;
      const float l = la * wa + lb * wb;
      const float inv = 1.0f / l;
      const int qi = t * 128 + row;
      if (qi < n) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float o0 = __uint_as_float(oa[j]) * wa, o1 = __uint_as_float(oa[j + 1]) * wa;
          if (nb > 0) {
            o0 = fmaf(__uint_as_float(ob[j]), wb, o0);
            o1 = fmaf(__uint_as_float(ob[j + 1]), wb, o1);
          }
          pk[j >> 1] = pack_bf16(o0 * inv, o1 * inv);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)(row0 + qi) * p.ldo + head * kHd);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        if (p.lse != nullptr)
          p.lse[((size_t)frame * p.heads + head) * n + qi] = m * p.scale_log2 + log2f(l);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace hma

extern "C" int hma_attn_spatial_fwd(const void* qkv, long long ld_qkv, int frames, int n, int heads, int q_col,
                                    int k_col, int v_col, float scale, void* out, long long ldo, float* lse,
                                    void* stream_) {
  using namespace hma;
  if (frames == 0) return 0;
  HMA_REQUIRE(n % 16 == 0 && n >= 16 && n <= kMaxN, "attn_spatial: tokens per frame n=%d must be a multiple of 16 in [16,320]", n);
  HMA_REQUIRE(heads >= 1, "attn_spatial: bad heads");
  AttnFwdParams p;
  p.n = n;
  p.na = n > 192 ? 192 : n;
  p.nb = n - p.na;
  p.box_rows = (n % 64 == 0) ? 64 : ((n % 32 == 0) ? 32 : 16);
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse = lse;
  p.heads = heads;
  CUtensorMap tm;
  int rc = hma_host::make_tmap_bf16_2d_sw(&tm, qkv, (uint64_t)ld_qkv, (uint64_t)frames * n, (uint64_t)ld_qkv * 2, 32,
                                          (uint32_t)p.box_rows, 64);
  if (rc) return rc;
  constexpr size_t smem = 1024 + 3 * kQKVBytes + kPBytes;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_spatial_fwd_kernel, dim3(frames * heads), dim3(kFwdThreads), smem,
                                      static_cast<cudaStream_t>(stream_), tm, p));
  return 0;
}
