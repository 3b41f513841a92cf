// Causal attention over the T frames of each spatial slot on the tensor cores
// (reference: attention.py:37-61 with causal=True, called from st_transformer.py:111).
//
// The sequences are tiny (T <= 128, head_dim 32), so SC = floor(128 / T) of them are packed into one
// 128-row tile: a 3-D TMA box {32 channels, SC slots, T frames} lifts q / k / v of one head for SC
// consecutive slots of one sample straight out of the (B, T, n, 3C) activation — row r of the tile is
// (frame r / SC, slot r % SC), so the "(B T) S C -> (B S) T C" transposes of the reference
// (st_transformer.py:89,113) never happen. S = Q K^T is one 128x128x32 tcgen05.mma; the softmax
// warps apply the block-diagonal causal mask (same slot, earlier-or-equal frame) from a per-row
// 128-bit mask, write P (bf16, 128-byte swizzle) and the P V product runs on the tensor core again.
// Seven eighths of S is masked away, which is irrelevant: the stage is bound by the bytes of qkv.
//
// The backward has the same shape as the spatial one with a single (key tile, query tile) pair.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct TemporalTcParams {
  int B, T, n, heads, SC;
  int q_col, k_col, v_col;
  float scale, scale_log2;
  __nv_bfloat16* out;           // fwd: [tokens, ldo]
  const __nv_bfloat16* out_c;   // bwd: forward output
  long long ldo;
  float* lse;                   // [tokens, heads], log2 domain (fwd: optional output, bwd: input)
  const __nv_bfloat16* dout;    // bwd
  long long ld_dout;
  __nv_bfloat16* dqkv;          // bwd
  long long ld_dqkv;
};

constexpr int kTRowB = 64;
constexpr int kTTile = 128 * kTRowB;   // 8 KB per operand tile
constexpr int kTPanel = 128 * 128;     // [128 x 64] bf16 panel

__device__ __forceinline__ uint64_t tdesc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(2048u >> 4) << 16;
  d |= static_cast<uint64_t>(512u >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;
  return d;
}

// 128-bit column mask of tile row r: bit c set iff column c is (same slot, frame <= frame of r)
__device__ __forceinline__ void row_mask(int r, int SC, int R, uint32_t (&m)[4]) {
  m[0] = m[1] = m[2] = m[3] = 0u;
  if (r >= R) return;
  const int sl = r % SC, t = r / SC;
  for (int tp = 0; tp <= t; ++tp) {
    const int c = tp * SC + sl;
    m[c >> 5] |= 1u << (c & 31);
  }
}

__device__ __forceinline__ void zero_tail_rows(uint32_t tile, int R, int tid, int nthreads) {
  // rows [R, 128) of a [128 x 64 B] tile are never written by the TMA box: clear them (16 B per store)
  const int vecs = (128 - R) * 4;
  for (int i = tid; i < vecs; i += nthreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile + (uint32_t)R * kTRowB + (uint32_t)i * 16), "r"(0u) : "memory");
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 4)
attn_temporal_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const TemporalTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sK = sQ + kTTile, sV = sK + kTTile, sP = sV + kTTile + 1024 * 0;  // 24 KB: aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (p.n + p.SC - 1) / p.SC;
  const int head = blockIdx.x % p.heads;
  const int chunk = (blockIdx.x / p.heads) % chunks;
  const int b = blockIdx.x / (p.heads * chunks);
  const int s0 = chunk * p.SC;
  const int R = p.SC * p.T;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_p), 128);
    mbar_init(smem_u32(&bar_o), 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&tmem_base_slot), 128);
    tmem_relinquish();
  }
  if (R < 128) {
    zero_tail_rows(sQ, R, threadIdx.x, blockDim.x);
    zero_tail_rows(sK, R, threadIdx.x, blockDim.x);
    zero_tail_rows(sV, R, threadIdx.x, blockDim.x);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t bl = smem_u32(&bar_load);
      mbar_expect_tx(bl, (uint32_t)(3 * R * kTRowB));
      tma_load_3d(sQ, &tmQKV, bl, p.q_col + head * 32, s0, b * p.T);
      tma_load_3d(sK, &tmQKV, bl, p.k_col + head * 32, s0, b * p.T);
      tma_load_3d(sV, &tmQKV, bl, p.v_col + head * 32, s0, b * p.T);
      mbar_wait(bl, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 32, 0, 1);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_ss(tmem_base, tdesc_sw64(sQ + k * 32), tdesc_sw64(sK + k * 32), idesc_s, (uint32_t)k);
      umma_commit(smem_u32(&bar_s));
      mbar_wait(smem_u32(&bar_p), 0);
      tc_fence_after();
      // O aliases the first 32 columns of S: every softmax thread has finished reading S (bar_p)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_ss(tmem_base, umma_desc_kmajor(sP + (uint32_t)(kk >> 2) * kTPanel + (uint32_t)(kk & 3) * 32),
                tdesc_sw64(sV + (uint32_t)kk * 16 * kTRowB), idesc_pv, (uint32_t)(kk != 0));
      umma_commit(smem_u32(&bar_o));
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t mask[4];
    row_mask(row, p.SC, R, mask);
    mbar_wait(smem_u32(&bar_s), 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_base + lane_addr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((mask[c] >> j) & 1u) m = fmaxf(m, __uint_as_float(r[j]));
    }
    const float mb = m * p.scale_log2;
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_base + lane_addr + c * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = ((mask[c] >> j) & 1u) ? fast_ex2(fmaf(__uint_as_float(r[j]), p.scale_log2, -mb)) : 0.f;
        const float p1 = ((mask[c] >> (j + 1)) & 1u) ? fast_ex2(fmaf(__uint_as_float(r[j + 1]), p.scale_log2, -mb)) : 0.f;
        const uint32_t w = pack_bf16(p0, p1);
        l += bf16_lo(w) + bf16_hi(w);
        pk[j >> 1] = w;
      }
      const uint32_t panel = sP + (uint32_t)(c >> 1) * kTPanel;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t addr = panel + sw128_offset((uint32_t)row, (uint32_t)((c & 1) * 32 + q * 8));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(smem_u32(&bar_p));
    mbar_wait(smem_u32(&bar_o), 0);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_x32(tmem_base + lane_addr, o);
    tmem_ld_wait();
    const int sl = row % p.SC, t = row / p.SC;
    if (row < R && s0 + sl < p.n) {
      const float inv = 1.0f / l;
      const size_t tok = ((size_t)b * p.T + t) * p.n + s0 + sl;
      uint4* dst = reinterpret_cast<uint4*>(p.out + tok * p.ldo + head * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
      if (p.lse != nullptr) p.lse[tok * p.heads + head] = mb + log2f(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const uint32_t (&r)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d4[q] = make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                       pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                       pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                       pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
}

__global__ void __launch_bounds__(288, 2)
attn_temporal_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                            const TemporalTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_sdp, bar_pds, bar_out;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sK = sQ + kTTile, sV = sK + kTTile, sDO = sV + kTTile;
  const uint32_t sP = sDO + kTTile;          // 32 KB offset: aligned
  const uint32_t sDS = sP + 2 * kTPanel;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (p.n + p.SC - 1) / p.SC;
  const int head = blockIdx.x % p.heads;
  const int chunk = (blockIdx.x / p.heads) % chunks;
  const int b = blockIdx.x / (p.heads * chunks);
  const int s0 = chunk * p.SC;
  const int R = p.SC * p.T;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_sdp), 1);
    mbar_init(smem_u32(&bar_pds), 256);
    mbar_init(smem_u32(&bar_out), 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(&tmem_base_slot), 256);
    tmem_relinquish();
  }
  if (R < 128) {
    zero_tail_rows(sQ, R, threadIdx.x, blockDim.x);
    zero_tail_rows(sK, R, threadIdx.x, blockDim.x);
    zero_tail_rows(sV, R, threadIdx.x, blockDim.x);
    zero_tail_rows(sDO, R, threadIdx.x, blockDim.x);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128;
  // after the compute warps have consumed S and dP their columns are reused for the three outputs
  const uint32_t tDV = tmem_base, tDK = tmem_base + 32, tDQ = tmem_base + 64;

  if (warp == 8) {
    if (elect_one()) {
      const uint32_t bl = smem_u32(&bar_load);
      mbar_expect_tx(bl, (uint32_t)(4 * R * kTRowB));
      tma_load_3d(sQ, &tmQKV, bl, p.q_col + head * 32, s0, b * p.T);
      tma_load_3d(sK, &tmQKV, bl, p.k_col + head * 32, s0, b * p.T);
      tma_load_3d(sV, &tmQKV, bl, p.v_col + head * 32, s0, b * p.T);
      tma_load_3d(sDO, &tmDO, bl, head * 32, s0, b * p.T);
      mbar_wait(bl, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_t = umma_idesc_bf16(128, 32, 1, 1);
      const uint32_t idesc_q = umma_idesc_bf16(128, 32, 0, 1);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_ss(tS, tdesc_sw64(sQ + k * 32), tdesc_sw64(sK + k * 32), idesc_s, (uint32_t)k);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_ss(tDP, tdesc_sw64(sDO + k * 32), tdesc_sw64(sV + k * 32), idesc_s, (uint32_t)k);
      umma_commit(smem_u32(&bar_sdp));
      mbar_wait(smem_u32(&bar_pds), 0);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        umma_ss(tDV, umma_desc_mnmajor(sP + kk * 2048, kTPanel), tdesc_sw64(sDO + kk * 1024), idesc_t, (uint32_t)(kk != 0));
        umma_ss(tDK, umma_desc_mnmajor(sDS + kk * 2048, kTPanel), tdesc_sw64(sQ + kk * 1024), idesc_t, (uint32_t)(kk != 0));
        umma_ss(tDQ, umma_desc_kmajor(sDS + (uint32_t)(kk >> 2) * kTPanel + (uint32_t)(kk & 3) * 32),
                tdesc_sw64(sK + (uint32_t)kk * 16 * kTRowB), idesc_q, (uint32_t)(kk != 0));
      }
      umma_commit(smem_u32(&bar_out));
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int sl = row % p.SC, t = row / p.SC;
    const bool valid = row < R && s0 + sl < p.n;
    const size_t tok = ((size_t)b * p.T + t) * p.n + s0 + sl;
    uint32_t mask[4];
    row_mask(valid ? row : 128, p.SC, R, mask);
    float L = 0.f, delta = 0.f;
    if (valid) {
      L = p.lse[tok * p.heads + head];
      const uint4* o4 = reinterpret_cast<const uint4*>(p.out_c + tok * p.ldo + head * 32);
      const uint4* g4 = reinterpret_cast<const uint4*>(p.dout + tok * p.ld_dout + head * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 a = o4[q], g = g4[q];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) delta += bf16_lo(aw[j]) * bf16_lo(gw[j]) + bf16_hi(aw[j]) * bf16_hi(gw[j]);
      }
    }
    mbar_wait(smem_u32(&bar_sdp), 0);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;  // 32-column chunk of the 128-key tile
      uint32_t s[32], dp[32];
      tmem_ld_x32(tS + lane_addr + c * 32, s);
      tmem_ld_x32(tDP + lane_addr + c * 32, dp);
      tmem_ld_wait();
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
        if ((mask[c] >> j) & 1u) {
          p0 = fast_ex2(fmaf(__uint_as_float(s[j]), p.scale_log2, -L));
          d0 = p0 * (__uint_as_float(dp[j]) - delta) * p.scale;
        }
        if ((mask[c] >> (j + 1)) & 1u) {
          p1 = fast_ex2(fmaf(__uint_as_float(s[j + 1]), p.scale_log2, -L));
          d1 = p1 * (__uint_as_float(dp[j + 1]) - delta) * p.scale;
        }
        pk[j >> 1] = pack_bf16(p0, p1);
        dk[j >> 1] = pack_bf16(d0, d1);
      }
      const uint32_t pan = (uint32_t)half * kTPanel;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = pan + sw128_offset((uint32_t)row, (uint32_t)(cc * 32 + q * 8));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(dk[4 * q]), "r"(dk[4 * q + 1]),
                     "r"(dk[4 * q + 2]), "r"(dk[4 * q + 3]) : "memory");
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(smem_u32(&bar_pds));
    mbar_wait(smem_u32(&bar_out), 0);
    tc_fence_after();
    // warps 0-3: dQ and dK of their rows; warps 4-7: dV
    if (half == 0) {
      uint32_t r[32];
      tmem_ld_x32(tDQ + lane_addr, r);
      tmem_ld_wait();
      if (valid) store_row32_bf16(p.dqkv + tok * p.ld_dqkv + p.q_col + head * 32, r);
      tmem_ld_x32(tDK + lane_addr, r);
      tmem_ld_wait();
      if (valid) store_row32_bf16(p.dqkv + tok * p.ld_dqkv + p.k_col + head * 32, r);
    } else {
      uint32_t r[32];
      tmem_ld_x32(tDV + lane_addr, r);
      tmem_ld_wait();
      if (valid) store_row32_bf16(p.dqkv + tok * p.ld_dqkv + p.v_col + head * 32, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int temporal_sc(int T) { return 128 / T; }

}  // namespace hma

extern "C" int hma_attn_temporal_fwd(const void* qkv, long long ld_qkv, int B, int T, int n, int heads, int q_col,
                                     int k_col, int v_col, float scale, void* out, long long ldo, float* lse,
                                     void* stream_) {
  using namespace hma;
  if (B == 0 || n == 0) return 0;
  HMA_REQUIRE(T >= 1 && T <= 128, "attn_temporal: T=%d must be in [1,128]", T);
  HMA_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
              "attn_temporal: 16-byte alignment required");
  TemporalTcParams p{};
  p.B = B; p.T = T; p.n = n; p.heads = heads; p.SC = temporal_sc(T);
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.lse = lse;
  CUtensorMap tm;
  int rc = hma_host::make_tmap_bf16_3d_sw64(&tm, qkv, (uint64_t)ld_qkv, (uint64_t)n, (uint64_t)B * T, (uint64_t)ld_qkv * 2,
                                            (uint64_t)n * ld_qkv * 2, (uint32_t)p.SC, (uint32_t)T);
  if (rc) return rc;
  constexpr size_t smem = 1024 + 3 * kTTile + 2 * kTPanel;
  static bool attr_done = false;
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int chunks = (n + p.SC - 1) / p.SC;
  attn_temporal_tc_fwd_kernel<<<B * chunks * heads, 160, smem, static_cast<cudaStream_t>(stream_)>>>(tm, p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_attn_temporal_bwd(const void* qkv, long long ld_qkv, const void* out, long long ldo,
                                     const void* dout, long long ld_dout, const float* lse, int B, int T, int n,
                                     int heads, int q_col, int k_col, int v_col, float scale, void* dqkv,
                                     long long ld_dqkv, void* stream_) {
  using namespace hma;
  if (B == 0 || n == 0) return 0;
  HMA_REQUIRE(T >= 1 && T <= 128, "attn_temporal_bwd: T=%d must be in [1,128]", T);
  HMA_REQUIRE(ld_qkv % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0 && ldo % 8 == 0, "attn_temporal_bwd: 16-byte alignment required");
  HMA_REQUIRE(lse != nullptr, "attn_temporal_bwd: needs the forward log-sum-exp");
  TemporalTcParams p{};
  p.B = B; p.T = T; p.n = n; p.heads = heads; p.SC = temporal_sc(T);
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.out_c = static_cast<const __nv_bfloat16*>(out); p.ldo = ldo; p.lse = const_cast<float*>(lse);
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.ld_dout = ld_dout;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  CUtensorMap tmQ, tmD;
  int rc = hma_host::make_tmap_bf16_3d_sw64(&tmQ, qkv, (uint64_t)ld_qkv, (uint64_t)n, (uint64_t)B * T, (uint64_t)ld_qkv * 2,
                                            (uint64_t)n * ld_qkv * 2, (uint32_t)p.SC, (uint32_t)T);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_3d_sw64(&tmD, dout, (uint64_t)ld_dout, (uint64_t)n, (uint64_t)B * T, (uint64_t)ld_dout * 2,
                                        (uint64_t)n * ld_dout * 2, (uint32_t)p.SC, (uint32_t)T);
  if (rc) return rc;
  constexpr size_t smem = 1024 + 4 * kTTile + 4 * kTPanel;
  static bool attr_done = false;
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int chunks = (n + p.SC - 1) / p.SC;
  attn_temporal_tc_bwd_kernel<<<B * chunks * heads, 288, smem, static_cast<cudaStream_t>(stream_)>>>(tmQ, tmD, p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
