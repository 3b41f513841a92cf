// Causal attention over the T frames of each spatial slot on the tensor cores
// (reference: attention.py:37-61 with causal=True, called from st_transformer.py:111), forward and
// backward, reading the (B, T, n, 3C) activation in place: the "(B T) S C -> (B S) T C" transposes of
// the reference (st_transformer.py:89,113) never happen.
//
// The sequences are tiny (T <= 128, head_dim 32) and the stage is bound by the bytes of qkv, so the
// design is about keeping HBM busy, not the tensor core:
//   * work unit = (sample, SC consecutive slots, head). Tp = T rounded up to a power of two, SC = 128 / Tp.
//     ONE 3-D TMA box {32 channels, Tp frames, SC slots} per operand lifts q / k / v of the unit out of the
//     activation as a 128-row tile in SLOT-MAJOR order (row = slot * Tp + frame): the score matrix
//     S = Q K^T (one 128x128x32 tcgen05.mma) is then block diagonal with Tp x Tp causal blocks, and a
//     softmax thread only touches the <= 32 columns of its own block (rows of padding frames / slots
//     beyond n are computed on zero-filled or foreign-but-finite data and never stored).
//   * persistent CTAs (one per SM) stream the units through a 6-stage (fwd) / 3-stage (bwd) TMA ring;
//     two softmax warpgroups alternate units, each with its own TMEM accumulators and P (/dS) tiles, and the
//     single UMMA-issuing thread polls its barriers, so the loads, the S contraction, the softmax, the P V
//     contraction and the stores of neighbouring units overlap. (Measured: the stage is bound by how many
//     units are in flight per SM -- a stage is occupied from TMA issue to the end of its P V -- not by
//     the softmax arithmetic.)
//   * P tiles are zeroed once: a thread always writes the same columns of its row, everything else
//     stays zero for the life of the CTA.
// Backward recomputes the softmax of the (<= 128-key) row from S, so it needs neither the forward
// output nor its log-sum-exp: delta = sum_j P_j dP_j is taken from the same registers.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct TemporalTcParams {
  int B, T, Tp, n, heads, SC;
  int chunks, units;
  int q_col, k_col, v_col;
  float scale, scale_log2;
  __nv_bfloat16* out;           // fwd: [tokens, ldo]
  long long ldo;
  float* lse;                   // fwd: optional [tokens, heads], log2 domain
  __nv_bfloat16* dqkv;          // bwd
  long long ld_dqkv;
};

constexpr int kTRowB = 64;
constexpr int kTTile = 128 * kTRowB;   // 8 KB per operand tile
constexpr int kTPanel = 128 * 128;     // [128 x 64] bf16 panel (16 KB); a P / dS tile is two panels
constexpr int kFwdStages = 6;   // units in flight per SM: a stage stays occupied from TMA issue until its P V completes
constexpr int kFwdGroups = 2;   // softmax warpgroups, one TMEM S buffer + one P tile each
constexpr int kBwdStages = 3;
constexpr int kGroupThreads = 128;
constexpr int kBwdThreads = 2 * kGroupThreads + 5 * 32;  // + producer, S/dP issuer, dV / dK / dQ issuers

__device__ __forceinline__ uint64_t tdesc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(2048u >> 4) << 16;
  d |= static_cast<uint64_t>(512u >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;
  return d;
}

// Descriptors are built once per role and advanced by adding (byte offset >> 4) to the low word: the single
// thread that issues the contractions is on the critical path of every unit.
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
constexpr uint32_t kHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO 512, version 1, SWIZZLE_64B
constexpr uint32_t kHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t lo_sw64(uint32_t addr) { return (addr >> 4) | ((2048u >> 4) << 16); }
__device__ __forceinline__ uint32_t lo_k128(uint32_t addr) { return (addr >> 4) | ((16u >> 4) << 16); }
__device__ __forceinline__ uint32_t lo_mn128(uint32_t addr) { return (addr >> 4) | ((uint32_t)(kTPanel >> 4) << 16); }

struct UnitCoord {
  int b, s0, head;
};
__device__ __forceinline__ UnitCoord unit_coord(const TemporalTcParams& p, int u) {
  UnitCoord c;
  c.head = u % p.heads;
  const int r = u / p.heads;
  c.s0 = (r % p.chunks) * p.SC;
  c.b = r / p.chunks;
  return c;
}

// What a softmax thread needs to know about its tile row.
struct RowInfo {
  int row, quarter;
  int sl, t;          // slot inside the unit, frame
  int off;            // first column of this row's block inside its 32-column chunk
  int nch;            // 32-column chunks per block (1 unless Tp > 32)
  uint32_t col0;      // first TMEM / P column of the row's chunk 0 (warp-uniform)
  bool frame_ok;      // t < T
};
__device__ __forceinline__ RowInfo row_info(const TemporalTcParams& p, int warp_in_group, int lane) {
  RowInfo r;
  r.quarter = warp_in_group;
  r.row = warp_in_group * 32 + lane;
  r.sl = r.row / p.Tp;
  r.t = r.row % p.Tp;
  if (p.Tp <= 32) {
    r.off = (lane / p.Tp) * p.Tp;
    r.nch = 1;
    r.col0 = (uint32_t)warp_in_group * 32u;
  } else {
    r.off = 0;
    r.nch = p.Tp / 32;
    r.col0 = (uint32_t)(r.sl * p.Tp);
  }
  r.frame_ok = r.t < p.T;
  return r;
}
// bit b of chunk j is set iff column (col0 + 32 j + b) is a key this row attends to (same slot, frame <= t)
__device__ __forceinline__ uint32_t chunk_mask(const RowInfo& r, int j) {
  if (!r.frame_ok) return 0u;  // padding rows attend to nothing: P = dS = 0, so they add nothing to dK / dV
  const int cnt = r.t - 32 * j + 1;
  if (cnt <= 0) return 0u;
  const uint32_t m = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
  return m << r.off;
}

__device__ __forceinline__ void store_p_chunk(uint32_t tile, int row, uint32_t col, const uint32_t (&pk)[16]) {
  const uint32_t panel = tile + (col >> 6) * kTPanel;
  const uint32_t c0 = col & 63u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = panel + sw128_offset((uint32_t)row, c0 + q * 8);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                 "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
  }
}

__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const uint32_t (&r)[32], float mul) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d4[q] = make_uint4(pack_bf16(__uint_as_float(r[8 * q]) * mul, __uint_as_float(r[8 * q + 1]) * mul),
                       pack_bf16(__uint_as_float(r[8 * q + 2]) * mul, __uint_as_float(r[8 * q + 3]) * mul),
                       pack_bf16(__uint_as_float(r[8 * q + 4]) * mul, __uint_as_float(r[8 * q + 5]) * mul),
                       pack_bf16(__uint_as_float(r[8 * q + 6]) * mul, __uint_as_float(r[8 * q + 7]) * mul));
}

__device__ __forceinline__ void zero_smem(uint32_t addr, int bytes, int tid, int nthreads) {
  for (int i = tid * 16; i < bytes; i += nthreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr + (uint32_t)i), "r"(0u) : "memory");
}

__device__ __forceinline__ float max32(const float (&s)[32]) {
  float a = fmaxf(s[0], s[1]), b = fmaxf(s[2], s[3]), c = fmaxf(s[4], s[5]), d = fmaxf(s[6], s[7]);
#pragma unroll
  for (int j = 8; j < 32; j += 4) {
    a = fmaxf(a, s[j]); b = fmaxf(b, s[j + 1]); c = fmaxf(c, s[j + 2]); d = fmaxf(d, s[j + 3]);
  }
  return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = kFwdGroups * kGroupThreads + 96;  // + producer, S issuer, P V issuer

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_temporal_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const TemporalTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kFwdStages], bar_empty[kFwdStages], bar_s[kFwdGroups], bar_p[kFwdGroups],
      bar_o[kFwdGroups], bar_ofree[kFwdGroups];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sStage = smem_base;                              // kFwdStages x {Q, K, V}
  const uint32_t sP = sStage + kFwdStages * 3 * kTTile;           // kFwdGroups x [128 x 128] bf16
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mine = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int kProducerWarp = kFwdGroups * 4, kMmaWarp = kProducerWarp + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int g = 0; g < kFwdGroups; ++g) {
      mbar_init(smem_u32(&bar_s[g]), 1);
      mbar_init(smem_u32(&bar_p[g]), kGroupThreads);
      mbar_init(smem_u32(&bar_o[g]), 1);
      mbar_init(smem_u32(&bar_ofree[g]), kGroupThreads);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  zero_smem(sP, kFwdGroups * 2 * kTPanel, threadIdx.x, blockDim.x);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == kProducerWarp) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQKV);
      for (int i = 0; i < n_mine; ++i) {
        const int st = i % kFwdStages;
        mbar_wait(smem_u32(&bar_empty[st]), (uint32_t)(((i / kFwdStages) & 1) ^ 1));
        const UnitCoord c = unit_coord(p, (int)blockIdx.x + i * (int)gridDim.x);
        const uint32_t full = smem_u32(&bar_full[st]);
        const uint32_t dst = sStage + (uint32_t)st * 3 * kTTile;
        mbar_expect_tx(full, 3u * kTTile);
        tma_load_3d(dst, &tmQKV, full, p.q_col + c.head * 32, c.b * p.T, c.s0);
        tma_load_3d(dst + kTTile, &tmQKV, full, p.k_col + c.head * 32, c.b * p.T, c.s0);
        tma_load_3d(dst + 2 * kTTile, &tmQKV, full, p.v_col + c.head * 32, c.b * p.T, c.s0);
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- S = Q K^T issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t q0 = lo_sw64(sStage);
      for (int is = 0; is < n_mine; ++is) {
        const int st = is % kFwdStages, g = is % kFwdGroups, k = is / kFwdGroups;
        mbar_wait(smem_u32(&bar_full[st]), (uint32_t)((is / kFwdStages) & 1));
        // S buffer g (which O aliases) is free once the epilogue of unit is - kFwdGroups has read O
        if (k >= 1) mbar_wait(smem_u32(&bar_ofree[g]), (uint32_t)((k - 1) & 1));
        tc_fence_after();
        const uint32_t qlo = q0 + (uint32_t)st * (3 * kTTile >> 4), klo = qlo + (kTTile >> 4);
        const uint32_t tS = tmem_base + (uint32_t)g * 128;
        umma_ss(tS, mk_desc(qlo, kHiSw64), mk_desc(klo, kHiSw64), idesc_s, 0u);
        umma_ss(tS, mk_desc(qlo + 2, kHiSw64), mk_desc(klo + 2, kHiSw64), idesc_s, 1u);
        umma_commit(smem_u32(&bar_s[g]));
      }
    }
  } else if (warp == kMmaWarp + 1) {
    // ---------------------------------------------------------------- O = P V issuer
    if (elect_one()) {
      const uint32_t idesc_pv = umma_idesc_bf16(128, 32, 0, 1);
      const uint32_t v0 = lo_sw64(sStage + 2 * kTTile), p0 = lo_k128(sP);
      for (int ip = 0; ip < n_mine; ++ip) {
        const int st = ip % kFwdStages, g = ip % kFwdGroups, k = ip / kFwdGroups;
        mbar_wait(smem_u32(&bar_p[g]), (uint32_t)(k & 1));
        tc_fence_after();
        const uint32_t plo = p0 + (uint32_t)g * (2 * kTPanel >> 4), vlo = v0 + (uint32_t)st * (3 * kTTile >> 4);
        const uint32_t tO = tmem_base + (uint32_t)g * 128;  // over the consumed S columns
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tO, mk_desc(plo + (uint32_t)(((kk >> 2) * kTPanel + (kk & 3) * 32) >> 4), kHiSw128),
                  mk_desc(vlo + (uint32_t)(kk * 1024 >> 4), kHiSw64), idesc_pv, (uint32_t)(kk != 0));
        umma_commit(smem_u32(&bar_o[g]));
        umma_commit(smem_u32(&bar_empty[st]));
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue, group g = warp / 4
    const int g = warp >> 2;
    const RowInfo ri = row_info(p, warp & 3, lane);
    const uint32_t lane_addr = (uint32_t)(ri.quarter * 32) << 16;
    const uint32_t tS = tmem_base + (uint32_t)g * 128 + lane_addr;
    const uint32_t tO = tS;
    const uint32_t tile = sP + (uint32_t)g * 2 * kTPanel;
    const uint32_t mk0 = chunk_mask(ri, 0);
    for (int i = g; i < n_mine; i += kFwdGroups) {
      const int k = i / kFwdGroups;
      const UnitCoord c = unit_coord(p, (int)blockIdx.x + i * (int)gridDim.x);
      mbar_wait(smem_u32(&bar_s[g]), (uint32_t)(k & 1));
      tc_fence_after();
      float l, mb;
      if (ri.nch == 1) {
        // whole block of the row in registers: mask folded into the scores, 4-way split reductions
        uint32_t r[32];
        tmem_ld_x32(tS + ri.col0, r);
        tmem_ld_wait();
        float sc[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) sc[b] = ((mk0 >> b) & 1u) ? __uint_as_float(r[b]) : -INFINITY;
        const float m = max32(sc);
        mb = ri.frame_ok ? m * p.scale_log2 : 0.f;
        uint32_t pk[16];
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 32; b += 2) {
          const uint32_t w = pack_bf16(fast_ex2(fmaf(sc[b], p.scale_log2, -mb)), fast_ex2(fmaf(sc[b + 1], p.scale_log2, -mb)));
          l4[(b >> 1) & 1] += bf16_lo(w);
          l4[2 + ((b >> 1) & 1)] += bf16_hi(w);
          pk[b >> 1] = w;
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        store_p_chunk(tile, ri.row, ri.col0, pk);
      } else {
        float m = -INFINITY;
        for (int j = 0; j < ri.nch; ++j) {
          uint32_t r[32];
          tmem_ld_x32(tS + ri.col0 + 32 * j, r);
          tmem_ld_wait();
          const uint32_t mk = chunk_mask(ri, j);
#pragma unroll
          for (int b = 0; b < 32; ++b)
            if ((mk >> b) & 1u) m = fmaxf(m, __uint_as_float(r[b]));
        }
        mb = m * p.scale_log2;
        l = 0.f;
        for (int j = 0; j < ri.nch; ++j) {
          uint32_t r[32];
          tmem_ld_x32(tS + ri.col0 + 32 * j, r);
          tmem_ld_wait();
          const uint32_t mk = chunk_mask(ri, j);
          uint32_t pk[16];
#pragma unroll
          for (int b = 0; b < 32; b += 2) {
            const float p0 = ((mk >> b) & 1u) ? fast_ex2(fmaf(__uint_as_float(r[b]), p.scale_log2, -mb)) : 0.f;
            const float p1 = ((mk >> (b + 1)) & 1u) ? fast_ex2(fmaf(__uint_as_float(r[b + 1]), p.scale_log2, -mb)) : 0.f;
            const uint32_t w = pack_bf16(p0, p1);
            l += bf16_lo(w) + bf16_hi(w);
            pk[b >> 1] = w;
          }
          store_p_chunk(tile, ri.row, ri.col0 + 32 * j, pk);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[g]));
      mbar_wait(smem_u32(&bar_o[g]), (uint32_t)(k & 1));
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_x32(tO, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_ofree[g]));
      if (ri.frame_ok && c.s0 + ri.sl < p.n) {
        const size_t tok = ((size_t)c.b * p.T + ri.t) * p.n + c.s0 + ri.sl;
        store_row32_bf16(p.out + tok * p.ldo + c.head * 32, o, 1.0f / l);
        if (p.lse != nullptr) p.lse[tok * p.heads + c.head] = mb + log2f(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_temporal_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                            const TemporalTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kBwdStages], bar_empty[kBwdStages], bar_sdp[2], bar_pds[2], bar_out[2], bar_free[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sStage = smem_base;                              // kBwdStages x {Q, K, V, dO}
  const uint32_t sPdS = sStage + kBwdStages * 4 * kTTile;         // 2 x {P, dS}, each [128 x 128] bf16
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mine = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 3);   // three output issuers release a stage
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(smem_u32(&bar_sdp[g]), 1);
      mbar_init(smem_u32(&bar_pds[g]), kGroupThreads);
      mbar_init(smem_u32(&bar_out[g]), 3);
      mbar_init(smem_u32(&bar_free[g]), kGroupThreads);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  zero_smem(sPdS, 8 * kTPanel, threadIdx.x, blockDim.x);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      for (int i = 0; i < n_mine; ++i) {
        const int st = i % kBwdStages;
        mbar_wait(smem_u32(&bar_empty[st]), (uint32_t)(((i / kBwdStages) & 1) ^ 1));
        const UnitCoord c = unit_coord(p, (int)blockIdx.x + i * (int)gridDim.x);
        const uint32_t full = smem_u32(&bar_full[st]);
        const uint32_t dst = sStage + (uint32_t)st * 4 * kTTile;
        mbar_expect_tx(full, 4u * kTTile);
        tma_load_3d(dst, &tmQKV, full, p.q_col + c.head * 32, c.b * p.T, c.s0);
        tma_load_3d(dst + kTTile, &tmQKV, full, p.k_col + c.head * 32, c.b * p.T, c.s0);
        tma_load_3d(dst + 2 * kTTile, &tmQKV, full, p.v_col + c.head * 32, c.b * p.T, c.s0);
        tma_load_3d(dst + 3 * kTTile, &tmDO, full, c.head * 32, c.b * p.T, c.s0);
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- S = Q K^T and dP = dO V^T issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t q0 = lo_sw64(sStage);
      for (int is = 0; is < n_mine; ++is) {
        const int st = is % kBwdStages, g = is & 1, k = is >> 1;
        mbar_wait(smem_u32(&bar_full[st]), (uint32_t)((is / kBwdStages) & 1));
        // the outputs of unit is-2 alias this S/dP buffer: wait until its epilogue has read them
        if (k >= 1) mbar_wait(smem_u32(&bar_free[g]), (uint32_t)((k - 1) & 1));
        tc_fence_after();
        const uint32_t qlo = q0 + (uint32_t)st * (4 * kTTile >> 4), klo = qlo + (kTTile >> 4), vlo = klo + (kTTile >> 4),
                       dlo = vlo + (kTTile >> 4);
        const uint32_t tS = tmem_base + (uint32_t)g * 256, tDP = tS + 128;
        umma_ss(tS, mk_desc(qlo, kHiSw64), mk_desc(klo, kHiSw64), idesc_s, 0u);
        umma_ss(tS, mk_desc(qlo + 2, kHiSw64), mk_desc(klo + 2, kHiSw64), idesc_s, 1u);
        umma_ss(tDP, mk_desc(dlo, kHiSw64), mk_desc(vlo, kHiSw64), idesc_s, 0u);
        umma_ss(tDP, mk_desc(dlo + 2, kHiSw64), mk_desc(vlo + 2, kHiSw64), idesc_s, 1u);
        umma_commit(smem_u32(&bar_sdp[g]));
      }
    }
  } else if (warp >= 10) {
    // ---------------------------------------------------------------- output issuers: warp 10 dV, 11 dK, 12 dQ
    if (elect_one()) {
      const int which = warp - 10;
      const uint32_t idesc = which == 2 ? umma_idesc_bf16(128, 32, 0, 1)   // dQ: A K-major, B MN-major
                                        : umma_idesc_bf16(128, 32, 1, 1);  // dV, dK: both operands MN-major
      // A: P (dV) or dS (dK, dQ) tile of the group; B: dO (dV), Q (dK) or K (dQ) tile of the stage
      const uint32_t a_base = sPdS + (which == 0 ? 0u : 2u * kTPanel);
      const uint32_t a0 = which == 2 ? lo_k128(a_base) : lo_mn128(a_base);
      const uint32_t b0 = lo_sw64(sStage + (which == 0 ? 3u : (which == 1 ? 0u : 1u)) * kTTile);
      const uint32_t tcol = which == 0 ? 0u : (which == 1 ? 32u : 64u);  // outputs reuse the consumed S / dP columns
      for (int ip = 0; ip < n_mine; ++ip) {
        const int st = ip % kBwdStages, g = ip & 1, k = ip >> 1;
        mbar_wait(smem_u32(&bar_pds[g]), (uint32_t)(k & 1));
        tc_fence_after();
        const uint32_t alo = a0 + (uint32_t)g * (4 * kTPanel >> 4), blo = b0 + (uint32_t)st * (4 * kTTile >> 4);
        const uint32_t tD = tmem_base + (uint32_t)g * 256 + tcol;
        if (which == 2) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tD, mk_desc(alo + (uint32_t)(((kk >> 2) * kTPanel + (kk & 3) * 32) >> 4), kHiSw128),
                    mk_desc(blo + (uint32_t)(kk * 1024 >> 4), kHiSw64), idesc, (uint32_t)(kk != 0));
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tD, mk_desc(alo + (uint32_t)(kk * 2048 >> 4), kHiSw128), mk_desc(blo + (uint32_t)(kk * 1024 >> 4), kHiSw64),
                    idesc, (uint32_t)(kk != 0));
        }
        umma_commit(smem_u32(&bar_out[g]));
        umma_commit(smem_u32(&bar_empty[st]));
      }
    }
  } else {
    const int g = warp >> 2;
    const RowInfo ri = row_info(p, warp & 3, lane);
    const uint32_t lane_addr = (uint32_t)(ri.quarter * 32) << 16;
    const uint32_t tS = tmem_base + (uint32_t)g * 256 + lane_addr, tDP = tS + 128;
    const uint32_t sPt = sPdS + (uint32_t)g * 4 * kTPanel, sDS = sPt + 2 * kTPanel;
    const uint32_t mk0 = chunk_mask(ri, 0);
    for (int i = g; i < n_mine; i += 2) {
      const int k = i >> 1;
      const UnitCoord c = unit_coord(p, (int)blockIdx.x + i * (int)gridDim.x);
      mbar_wait(smem_u32(&bar_sdp[g]), (uint32_t)(k & 1));
      tc_fence_after();
      if (ri.nch == 1) {
        // the whole block of the row lives in registers: one TMEM read of S and dP
        uint32_t sr[32], dpr[32];
        tmem_ld_x32(tS + ri.col0, sr);
        tmem_ld_x32(tDP + ri.col0, dpr);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) e[b] = ((mk0 >> b) & 1u) ? __uint_as_float(sr[b]) : -INFINITY;
        const float m = max32(e);
        const float mb = ri.frame_ok ? m * p.scale_log2 : 0.f;
        float l4[4] = {0.f, 0.f, 0.f, 0.f}, d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 32; ++b) {
          e[b] = fast_ex2(fmaf(e[b], p.scale_log2, -mb));
          l4[b & 3] += e[b];
          d4[b & 3] = fmaf(e[b], __uint_as_float(dpr[b]), d4[b & 3]);
        }
        const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        const float inv = l > 0.f ? 1.0f / l : 0.f;  // l == 0 only on padding rows
        const float delta = ((d4[0] + d4[1]) + (d4[2] + d4[3])) * inv;
        const float inv_s = inv * p.scale;
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int b = 0; b < 32; b += 2) {
          pk[b >> 1] = pack_bf16(e[b] * inv, e[b + 1] * inv);
          dk[b >> 1] = pack_bf16(e[b] * inv_s * (__uint_as_float(dpr[b]) - delta),
                                 e[b + 1] * inv_s * (__uint_as_float(dpr[b + 1]) - delta));
        }
        store_p_chunk(sPt, ri.row, ri.col0, pk);
        store_p_chunk(sDS, ri.row, ri.col0, dk);
      } else {
        // pass 1: row maximum
        float m = -INFINITY;
        for (int j = 0; j < ri.nch; ++j) {
          uint32_t s[32];
          tmem_ld_x32(tS + ri.col0 + 32 * j, s);
          tmem_ld_wait();
          const uint32_t mk = chunk_mask(ri, j);
#pragma unroll
          for (int b = 0; b < 32; ++b)
            if ((mk >> b) & 1u) m = fmaxf(m, __uint_as_float(s[b]));
        }
        const float mb = m * p.scale_log2;
        // pass 2: softmax denominator and delta = sum_j P_j dP_j
        float l = 0.f, pd = 0.f;
        for (int j = 0; j < ri.nch; ++j) {
          uint32_t s[32], dp[32];
          tmem_ld_x32(tS + ri.col0 + 32 * j, s);
          tmem_ld_x32(tDP + ri.col0 + 32 * j, dp);
          tmem_ld_wait();
          const uint32_t mk = chunk_mask(ri, j);
#pragma unroll
          for (int b = 0; b < 32; ++b) {
            const float e = ((mk >> b) & 1u) ? fast_ex2(fmaf(__uint_as_float(s[b]), p.scale_log2, -mb)) : 0.f;
            l += e;
            pd = fmaf(e, __uint_as_float(dp[b]), pd);
          }
        }
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        const float delta = pd * inv;
        // pass 3: P and dS tiles
        for (int j = 0; j < ri.nch; ++j) {
          uint32_t s[32], dp[32];
          tmem_ld_x32(tS + ri.col0 + 32 * j, s);
          tmem_ld_x32(tDP + ri.col0 + 32 * j, dp);
          tmem_ld_wait();
          const uint32_t mk = chunk_mask(ri, j);
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int b = 0; b < 32; b += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if ((mk >> b) & 1u) {
              p0 = fast_ex2(fmaf(__uint_as_float(s[b]), p.scale_log2, -mb)) * inv;
              d0 = p0 * (__uint_as_float(dp[b]) - delta) * p.scale;
            }
            if ((mk >> (b + 1)) & 1u) {
              p1 = fast_ex2(fmaf(__uint_as_float(s[b + 1]), p.scale_log2, -mb)) * inv;
              d1 = p1 * (__uint_as_float(dp[b + 1]) - delta) * p.scale;
            }
            pk[b >> 1] = pack_bf16(p0, p1);
            dk[b >> 1] = pack_bf16(d0, d1);
          }
          store_p_chunk(sPt, ri.row, ri.col0 + 32 * j, pk);
          store_p_chunk(sDS, ri.row, ri.col0 + 32 * j, dk);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_pds[g]));
      mbar_wait(smem_u32(&bar_out[g]), (uint32_t)(k & 1));
      tc_fence_after();
      const uint32_t tb = tmem_base + (uint32_t)g * 256 + lane_addr;
      const bool valid = ri.frame_ok && c.s0 + ri.sl < p.n;
      const size_t tok = ((size_t)c.b * p.T + ri.t) * p.n + c.s0 + ri.sl;
      __nv_bfloat16* drow = p.dqkv + (valid ? tok * p.ld_dqkv : 0) + c.head * 32;
      uint32_t rq[32], rk[32], rv[32];
      tmem_ld_x32(tb + 64, rq);
      tmem_ld_x32(tb + 32, rk);
      tmem_ld_x32(tb, rv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_free[g]));
      if (valid) {
        store_row32_bf16(drow + p.q_col, rq, 1.0f);
        store_row32_bf16(drow + p.k_col, rk, 1.0f);
        store_row32_bf16(drow + p.v_col, rv, 1.0f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int fill_params(TemporalTcParams& p, int B, int T, int n, int heads, int q_col, int k_col, int v_col, float scale) {
  HMA_REQUIRE(T >= 1 && T <= 128, "attn_temporal: T=%d must be in [1,128]", T);
  HMA_REQUIRE(heads >= 1, "attn_temporal: bad heads");
  p.B = B; p.T = T; p.n = n; p.heads = heads;
  int tp = 1;
  while (tp < T) tp <<= 1;
  p.Tp = tp;
  p.SC = 128 / tp;
  p.chunks = (n + p.SC - 1) / p.SC;
  p.units = B * p.chunks * heads;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  return 0;
}

// {32 channels, Tp frames, SC slots} box over the (B*T, n, ld) activation: dimension 1 = frames (stride n*ld),
// dimension 2 = slots (stride ld), so the tile lands slot-major in shared memory.
static int make_unit_map(CUtensorMap* tm, const void* base, long long ld, int B, int T, int n, const TemporalTcParams& p) {
  return hma_host::make_tmap_bf16_3d_sw64(tm, base, (uint64_t)ld, (uint64_t)B * T, (uint64_t)n, (uint64_t)n * ld * 2,
                                          (uint64_t)ld * 2, (uint32_t)p.Tp, (uint32_t)p.SC);
}

}  // namespace hma

extern "C" int hma_attn_temporal_fwd(const void* qkv, long long ld_qkv, int B, int T, int n, int heads, int q_col,
                                     int k_col, int v_col, float scale, void* out, long long ldo, float* lse,
                                     void* stream_) {
  using namespace hma;
  if (B == 0 || n == 0) return 0;
  HMA_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
              "attn_temporal: 16-byte alignment required");
  TemporalTcParams p{};
  if (int rc = fill_params(p, B, T, n, heads, q_col, k_col, v_col, scale)) return rc;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.lse = lse;
  CUtensorMap tm;
  if (int rc = make_unit_map(&tm, qkv, ld_qkv, B, T, n, p)) return rc;
  constexpr size_t smem = 1024 + kFwdStages * 3 * kTTile + kFwdGroups * 2 * kTPanel;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int grid = hma_host::sm_count();
  if (grid > p.units) grid = p.units;
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_temporal_tc_fwd_kernel, dim3(grid), dim3(kFwdThreads), smem,
                                      static_cast<cudaStream_t>(stream_), tm, p));
  return 0;
}

// `out` and `lse` of the forward are accepted for interface symmetry with the spatial kernel but not read:
// the backward recomputes the (tiny) softmax rows.
extern "C" int hma_attn_temporal_bwd(const void* qkv, long long ld_qkv, const void* out, long long ldo,
                                     const void* dout, long long ld_dout, const float* lse, int B, int T, int n,
                                     int heads, int q_col, int k_col, int v_col, float scale, void* dqkv,
                                     long long ld_dqkv, void* stream_) {
  using namespace hma;
  (void)out; (void)ldo; (void)lse;
  if (B == 0 || n == 0) return 0;
  HMA_REQUIRE(ld_qkv % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "attn_temporal_bwd: 16-byte alignment required");
  HMA_REQUIRE(q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, "attn_temporal_bwd: 16-byte alignment required");
  TemporalTcParams p{};
  if (int rc = fill_params(p, B, T, n, heads, q_col, k_col, v_col, scale)) return rc;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  CUtensorMap tmQ, tmD;
  if (int rc = make_unit_map(&tmQ, qkv, ld_qkv, B, T, n, p)) return rc;
  if (int rc = make_unit_map(&tmD, dout, ld_dout, B, T, n, p)) return rc;
  constexpr size_t smem = 1024 + kBwdStages * 4 * kTTile + 8 * kTPanel;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int grid = hma_host::sm_count();
  if (grid > p.units) grid = p.units;
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_temporal_tc_bwd_kernel, dim3(grid), dim3(kBwdThreads), smem,
                                      static_cast<cudaStream_t>(stream_), tmQ, tmD, p));
  return 0;
}
