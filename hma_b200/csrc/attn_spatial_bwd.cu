// Backward of the per-frame bidirectional attention (autograd transpose of attention.py:37-61 as
// called from st_transformer.py:85-86). With P = softmax(scale * Q K^T) recomputed from the saved
// log-sum-exp:
//   dV = P^T dO        dP = dO V^T        dS = P * (dP - rowsum(dO * O)) * scale
//   dQ = dS K          dK = dS^T Q
// One CTA per (frame, head); five tcgen05 contractions per (key tile, query tile) pair of 128 x 128:
//   S, dP        : K-major operands straight from the 64B-swizzled q/k/v/dO tiles TMA loaded
//   dV, dK       : A = P^T / dS^T read MN-major from the bf16 tiles the compute warps wrote,
//                  B = dO / Q read MN-major (reduction over query rows)
//   dQ           : A = dS K-major, B = K MN-major (reduction over keys)
// Accumulators live in TMEM (S 128 + dP 128 + dQ 3x32 + 2 x (dK 32 + dV 32) columns).
//
// The element-wise work is a two-stage, warp-specialised pipeline (both stages: warp w owns TMEM lane quarter w%4
// and a 64-key column half):
//   stage A (warps 0-7)   S -> P = exp2(scale*S - lse) as bf16 into shared memory        (MUFU-bound)
//   stage B (warps 8-15)  dP, P (read back from shared memory) -> dS = P*(dP - delta)*scale  (FMA-bound)
// so the exp phase of tile pair i+1 overlaps the dS phase of pair i on every SM sub-partition instead of all
// sixteen warps sitting in the same phase, S and dP are handed back to the tensor core separately (S as soon as
// stage A has copied it out, dP as soon as stage B has), and dV = P^T dO is issued as soon as P exists, before dS.
// The P / dS tiles are double buffered in shared memory so neither stage waits for the dV / dK / dQ contractions
// of the previous pair. Operand tiles are loaded per 128-row group with their own barrier: the first pair starts
// after 32 KB of the CTA's 80 KB have landed.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct AttnBwdParams {
  int n, box_rows, heads;
  int q_col, k_col, v_col;
  float scale, scale_log2;
  const __nv_bfloat16* out;   // forward output, [tokens, ldo]
  long long ldo;
  const __nv_bfloat16* dout;  // [tokens, ld_dout]
  long long ld_dout;
  const float* lse;           // [frames, heads, n]
  const float* delta;         // [tokens, heads] = rowsum(dO * O) per (token, head), or null: computed here from out / dout
  __nv_bfloat16* dqkv;        // [tokens, ld_dqkv]
  long long ld_dqkv;
};

constexpr int kBRowB = 64;
constexpr int kBMaxN = 320;
constexpr int kBTile = kBMaxN * kBRowB;   // 20 KB per operand
constexpr int kBPanel = 128 * 128;        // [128 x 64] bf16 panel
constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kStageThreads = kComputeThreads / 2;  // threads of one pipeline stage (A: S -> P, B: dP -> dS)
constexpr int kBwdThreads = kComputeThreads + 128;  // + four UMMA-issuing warps (the first also issues the TMA loads)

__device__ __forceinline__ uint64_t bdesc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(2048u >> 4) << 16;
  d |= static_cast<uint64_t>(512u >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;
  return d;
}

__device__ __forceinline__ void store_head_row(__nv_bfloat16* dst, const uint32_t (&r)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d4[q] = make_uint4(pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                       pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                       pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                       pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_spatial_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load[3], bar_s, bar_dp, bar_sfree, bar_dpfree, bar_p[2], bar_ds[2], bar_pfree[2], bar_dsfree[2], bar_kv,
      bar_epi[2], bar_final;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_lse[kBMaxN];
  __shared__ float s_delta[kBMaxN];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + kBTile;
  const uint32_t sV = sK + kBTile;
  const uint32_t sDO = sV + kBTile;
  const uint32_t sP0 = sDO + kBTile;         // 80 KB offset: 1024-aligned; two {P, dS} buffers of 64 KB
  constexpr uint32_t kPdsBuf = 4 * kBPanel;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int frame = blockIdx.x / p.heads;
  const int head = blockIdx.x % p.heads;
  const int n = p.n;
  const int ntile = (n + 127) / 128;
  const int row0 = frame * n;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bar_load[i]), 1);
    mbar_init(smem_u32(&bar_s), 1);
    mbar_init(smem_u32(&bar_dp), 1);
    mbar_init(smem_u32(&bar_sfree), kStageThreads);
    mbar_init(smem_u32(&bar_dpfree), kStageThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_p[i]), kStageThreads);
      mbar_init(smem_u32(&bar_ds[i]), kStageThreads);
      mbar_init(smem_u32(&bar_pfree[i]), 1);   // dV issued from P[i] has completed: stage A may overwrite P[i]
      mbar_init(smem_u32(&bar_dsfree[i]), 2);  // dK and dQ have consumed dS[i]
    }
    mbar_init(smem_u32(&bar_kv), 2);  // dV issuer + dK issuer: both accumulators of the key tile are final
    mbar_init(smem_u32(&bar_epi[0]), kStageThreads);
    mbar_init(smem_u32(&bar_epi[1]), kStageThreads);
    mbar_init(smem_u32(&bar_final), 1);
    fence_barrier_init();
  }
  if (warp == kComputeWarps) {  // TMEM allocation overlaps the predecessor's tail (before griddepcontrol.wait)
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) HMA_TL(0, 0);
  if (warp == kComputeWarps) {
    // start the operand loads first: they overlap the per-row statistics below. One barrier per 128-row group of
    // all four operands, in the order the tile pairs need them.
    if (elect_one()) {
      for (int g = 0; g < ntile; ++g) {
        const int r_lo = g * 128, r_hi = min(n, r_lo + 128);
        const uint32_t bl = smem_u32(&bar_load[g]);
        mbar_expect_tx(bl, (uint32_t)(4 * (r_hi - r_lo) * kBRowB));
        for (int r = r_lo; r < r_hi; r += p.box_rows) {
          tma_load_2d(sQ + r * kBRowB, &tmQKV, bl, p.q_col + head * 32, row0 + r);
          tma_load_2d(sK + r * kBRowB, &tmQKV, bl, p.k_col + head * 32, row0 + r);
          tma_load_2d(sV + r * kBRowB, &tmQKV, bl, p.v_col + head * 32, row0 + r);
          tma_load_2d(sDO + r * kBRowB, &tmDO, bl, head * 32, row0 + r);
        }
      }
    }
    __syncwarp();
    if (lane == 0) HMA_TL(0, 1);
  }
  // per-row statistics: log2-sum-exp and delta = dO . O (the latter normally arrives precomputed from the epilogue of the
  // GEMM that produced dO, so that the prologue only reads 8 bytes per row)
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    s_lse[r] = p.lse[((size_t)frame * p.heads + head) * n + r];
    if (p.delta != nullptr) {
      s_delta[r] = p.delta[(size_t)(row0 + r) * p.heads + head];
      continue;
    }
    const uint4* o4 = reinterpret_cast<const uint4*>(p.out + (size_t)(row0 + r) * p.ldo + head * 32);
    const uint4* g4 = reinterpret_cast<const uint4*>(p.dout + (size_t)(row0 + r) * p.ld_dout + head * 32);
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 a = o4[q], b = g4[q];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += bf16_lo(aw[j]) * bf16_lo(bw[j]) + bf16_hi(aw[j]) * bf16_hi(bw[j]);
    }
    s_delta[r] = acc;
  }
  if (threadIdx.x == 0) HMA_TL(0, 3);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) HMA_TL(0, 4);
  const uint32_t tmem_base = tmem_base_slot;
  // dK / dV accumulators are double-buffered over key tiles (buffer kt & 1, kAccBuf columns apart): the contractions of
  // key tile kt + 1 start while tile kt's accumulators are still waiting to be written out
  constexpr uint32_t kAccBuf = 64;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDQ = tmem_base + 256, tDK = tmem_base + 352,
                 tDV = tmem_base + 384;

  if (warp >= kComputeWarps) {
    // ---------------------------------------------------------------- four UMMA issuers (one elected lane each):
    //   warp 16: S = Q K^T of the next tile pair as soon as stage A has copied S out, dP = dO V^T as soon as stage B
    //            has copied dP out
    //   warp 17: dV += P^T dO (needs P only)     warp 18: dK += dS^T Q     warp 19: dQ += dS K
    // An issuing thread competes for issue slots with the four or five compute warps of its scheduler and gets one every
    // few tens of cycles: the clock64 timelines showed ONE thread issuing both dV and dK (16 tcgen05.mma and their
    // descriptor arithmetic per tile pair) setting the period of the whole pipeline, so every contraction has its own
    // issuer and the issue loops carry no per-instruction predicates. Descriptors are a constant high word plus a low word
    // advanced by (byte offset >> 4).
    if (elect_one()) {
      const int role = warp - kComputeWarps;
      const int NI = ntile * ntile;
      constexpr uint32_t kHi64 = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO 512, version 1, SWIZZLE_64B
      constexpr uint32_t kHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
      auto lo64 = [](uint32_t addr) { return (addr >> 4) | ((2048u >> 4) << 16); };
      auto mk = [](uint32_t lo, uint32_t hi) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
        return d;
      };
      if (role == 0) {
        const uint32_t q0 = lo64(sQ), k0 = lo64(sK), v0 = lo64(sV), d0 = lo64(sDO);
        constexpr uint32_t kTileStep = (128u * kBRowB) >> 4;
        int kt = 0, qt = 0;
        for (int it = 0; it < NI; ++it) {
          // operand groups of this pair (a completed barrier stays passable: its phase is never re-armed)
          mbar_wait(smem_u32(&bar_load[kt]), 0);
          mbar_wait(smem_u32(&bar_load[qt]), 0);
          if (it == 0) HMA_TL(11, 0);
          const int nk = min(128, n - kt * 128);
          const uint32_t idesc = umma_idesc_bf16(128, nk, 0, 0);
          const uint32_t ql = q0 + (uint32_t)qt * kTileStep, dl = d0 + (uint32_t)qt * kTileStep;
          const uint32_t kl = k0 + (uint32_t)kt * kTileStep, vl = v0 + (uint32_t)kt * kTileStep;
          if (it > 0) mbar_wait(smem_u32(&bar_sfree), (uint32_t)((it - 1) & 1));
          tc_fence_after();
          umma_ss(tS, mk(ql, kHi64), mk(kl, kHi64), idesc, 0u);
          umma_ss(tS, mk(ql + 2, kHi64), mk(kl + 2, kHi64), idesc, 1u);
          umma_commit(smem_u32(&bar_s));
          if (it > 0) {
            mbar_wait(smem_u32(&bar_dpfree), (uint32_t)((it - 1) & 1));
            tc_fence_after();
          }
          umma_ss(tDP, mk(dl, kHi64), mk(vl, kHi64), idesc, 0u);
          umma_ss(tDP, mk(dl + 2, kHi64), mk(vl + 2, kHi64), idesc, 1u);
          umma_commit(smem_u32(&bar_dp));
          HMA_TL(1, it);
          if (++qt == ntile) { qt = 0; ++kt; }
        }
      } else if (role == 1) {
        // ------------------------------------------------ dV += P^T dO: needs P only, so it runs ahead of dS
        const uint32_t idesc_t = umma_idesc_bf16(128, 32, 1, 1);   // both operands MN-major
        const uint32_t p0 = (sP0 >> 4) | ((uint32_t)(kBPanel >> 4) << 16);
        const uint32_t d0 = lo64(sDO);
        int kt = 0, qt = 0;
        for (int it = 0; it < NI; ++it) {
          const int bsel = it & 1;
          const int kq16 = min(128, n - qt * 128) >> 4;
          const uint32_t pl = p0 + (uint32_t)bsel * (kPdsBuf >> 4);
          const uint32_t dl = d0 + (uint32_t)qt * ((128u * kBRowB) >> 4);
          const uint32_t acc0 = (uint32_t)(qt != 0);
          mbar_wait(smem_u32(&bar_load[qt]), 0);  // dO tile qt (long since landed: S of this pair needed it too)
          mbar_wait(smem_u32(&bar_p[bsel]), (uint32_t)((it >> 1) & 1));
          HMA_TL(2, it);
          const uint32_t tD = tDV + (uint32_t)(kt & 1) * kAccBuf;
          if (qt == 0 && kt >= 2) {  // the dV that last used this accumulator buffer has been read out of TMEM
            mbar_wait(smem_u32(&bar_epi[kt & 1]), (uint32_t)(((kt >> 1) - 1) & 1));
            tc_fence_after();
          }
          if (kq16 == 8) {
            umma_ss(tD, mk(pl, kHi128), mk(dl, kHi64), idesc_t, acc0);
#pragma unroll
            for (int kk = 1; kk < 8; ++kk)
              umma_ss(tD, mk(pl + (uint32_t)(kk * 2048 >> 4), kHi128), mk(dl + (uint32_t)(kk * 1024 >> 4), kHi64), idesc_t, 1u);
          } else {
            for (int kk = 0; kk < kq16; ++kk)
              umma_ss(tD, mk(pl + (uint32_t)kk * (2048u >> 4), kHi128), mk(dl + (uint32_t)kk * (1024u >> 4), kHi64), idesc_t,
                      kk == 0 ? acc0 : 1u);
          }
          umma_commit(smem_u32(&bar_pfree[bsel]));
          if (qt == ntile - 1) umma_commit(smem_u32(&bar_kv));
          HMA_TL(3, it);
          if (++qt == ntile) { qt = 0; ++kt; }
        }
      } else if (role == 2) {
        // ------------------------------------------------ dK += dS^T Q
        const uint32_t idesc_t = umma_idesc_bf16(128, 32, 1, 1);
        const uint32_t s0 = ((sP0 + 2 * kBPanel) >> 4) | ((uint32_t)(kBPanel >> 4) << 16);
        const uint32_t q0 = lo64(sQ);
        int kt = 0, qt = 0;
        for (int it = 0; it < NI; ++it) {
          const int bsel = it & 1;
          const int kq16 = min(128, n - qt * 128) >> 4;
          const uint32_t sl = s0 + (uint32_t)bsel * (kPdsBuf >> 4);
          const uint32_t ql = q0 + (uint32_t)qt * ((128u * kBRowB) >> 4);
          const uint32_t acc0 = (uint32_t)(qt != 0);
          mbar_wait(smem_u32(&bar_load[qt]), 0);
          mbar_wait(smem_u32(&bar_ds[bsel]), (uint32_t)((it >> 1) & 1));
          const uint32_t tD = tDK + (uint32_t)(kt & 1) * kAccBuf;
          if (qt == 0 && kt >= 2) {  // the dK that last used this accumulator buffer has been read out of TMEM
            mbar_wait(smem_u32(&bar_epi[kt & 1]), (uint32_t)(((kt >> 1) - 1) & 1));
            tc_fence_after();
          }
          if (kq16 == 8) {
            umma_ss(tD, mk(sl, kHi128), mk(ql, kHi64), idesc_t, acc0);
#pragma unroll
            for (int kk = 1; kk < 8; ++kk)
              umma_ss(tD, mk(sl + (uint32_t)(kk * 2048 >> 4), kHi128), mk(ql + (uint32_t)(kk * 1024 >> 4), kHi64), idesc_t, 1u);
          } else {
            for (int kk = 0; kk < kq16; ++kk)
              umma_ss(tD, mk(sl + (uint32_t)kk * (2048u >> 4), kHi128), mk(ql + (uint32_t)kk * (1024u >> 4), kHi64), idesc_t,
                      kk == 0 ? acc0 : 1u);
          }
          umma_commit(smem_u32(&bar_dsfree[bsel]));
          if (qt == ntile - 1) umma_commit(smem_u32(&bar_kv));
          HMA_TL(4, it);
          if (++qt == ntile) { qt = 0; ++kt; }
        }
      } else {
        // ------------------------------------------------ dQ += dS K
        const uint32_t idesc_q = umma_idesc_bf16(128, 32, 0, 1);   // A K-major, B MN-major
        const uint32_t s0 = ((sP0 + 2 * kBPanel) >> 4) | ((16u >> 4) << 16);
        const uint32_t k0 = lo64(sK);
        int kt = 0, qt = 0;
        for (int it = 0; it < NI; ++it) {
          const int bsel = it & 1;
          const int nk16 = min(128, n - kt * 128) >> 4;
          const uint32_t sl = s0 + (uint32_t)bsel * (kPdsBuf >> 4);
          const uint32_t kl = k0 + (uint32_t)kt * ((128u * kBRowB) >> 4);
          const uint32_t tD = tDQ + (uint32_t)qt * 32;
          const uint32_t acc0 = (uint32_t)(kt != 0);
          mbar_wait(smem_u32(&bar_load[kt]), 0);
          mbar_wait(smem_u32(&bar_ds[bsel]), (uint32_t)((it >> 1) & 1));
          if (nk16 == 8) {
            umma_ss(tD, mk(sl, kHi128), mk(kl, kHi64), idesc_q, acc0);
#pragma unroll
            for (int kk = 1; kk < 8; ++kk)
              umma_ss(tD, mk(sl + (uint32_t)(((kk >> 2) * kBPanel + (kk & 3) * 32) >> 4), kHi128),
                      mk(kl + (uint32_t)(kk * 16 * kBRowB >> 4), kHi64), idesc_q, 1u);
          } else {
            for (int kk = 0; kk < nk16; ++kk)
              umma_ss(tD, mk(sl + (uint32_t)(((kk >> 2) * kBPanel + (kk & 3) * 32) >> 4), kHi128),
                      mk(kl + (uint32_t)kk * (uint32_t)(16 * kBRowB >> 4), kHi64), idesc_q, kk == 0 ? acc0 : 1u);
          }
          umma_commit(smem_u32(&bar_dsfree[bsel]));
          if (it == NI - 1) umma_commit(smem_u32(&bar_final));
          if (++qt == ntile) { qt = 0; ++kt; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps 0..15
    const bool stage_b = warp >= 8;
    const int quarter = warp & 3;           // TMEM lane quarter (query rows)
    const int half = (warp >> 2) & 1;       // key columns [64*half, 64*half + 64) of the 128-key tile
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t pan = (uint32_t)half * kBPanel;  // a 64-column half is exactly one 128-byte-swizzled panel
    if (!stage_b) {
      // ------------------------------------------------ stage A: S -> P
      int it = 0;
      for (int kt = 0; kt < ntile; ++kt) {
        const int nk = min(128, n - kt * 128);
        for (int qt = 0; qt < ntile; ++qt, ++it) {
          const int bsel = it & 1;
          if (it >= 2) {  // P[bsel] of pair it-2 has been consumed by dV (tensor core) and by stage B (its dS is out)
            mbar_wait(smem_u32(&bar_pfree[bsel]), (uint32_t)(((it - 2) >> 1) & 1));
            mbar_wait(smem_u32(&bar_ds[bsel]), (uint32_t)(((it - 2) >> 1) & 1));
          }
          mbar_wait(smem_u32(&bar_s), (uint32_t)(it & 1));
          tc_fence_after();
          const int qi = qt * 128 + row;
          const float L = qi < n ? s_lse[qi] : 0.f;
          // warp-uniform: rows past the frame are never read by the dV / dK contractions and only produce dQ rows
          // that are not stored; key columns past nk are never read at all
          const bool rows_live = quarter * 32 < min(128, n - qt * 128);
          const uint32_t sP = sP0 + (uint32_t)bsel * kPdsBuf + pan;
          // both 32-column chunks of this thread's half go to registers first, so S returns to the tensor core at once
          // and the next pair's S = Q K^T is computed while this pair's exponentials are
          uint32_t s0[32], s1[32];
          tmem_ld_x32(tS + lane_addr + (uint32_t)(half * 64), s0);  // unconditional: allocated columns
          tmem_ld_x32(tS + lane_addr + (uint32_t)(half * 64 + 32), s1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&bar_sfree));
          auto exp_chunk = [&](const uint32_t (&sv)[32], int c) {
            if (rows_live && half * 64 + c * 32 < nk) {
              uint32_t pk[16];
              const f32x2 sc2 = dup2(p.scale_log2), nl2 = dup2(-L);
#pragma unroll
              for (int j = 0; j < 32; j += 2) {  // one packed FFMA2 per two scores (issue slots are what this stage is short of)
                float a, b;
                upk2(fma2(pk2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), sc2, nl2), a, b);
                pk[j >> 1] = pack_bf16(fast_ex2(a), fast_ex2(b));
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(c * 32 + q * 8));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                             "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
              }
            }
          };
          exp_chunk(s0, 0);
          exp_chunk(s1, 1);
          fence_proxy_async();
          mbar_arrive(smem_u32(&bar_p[bsel]));
          if (lane == 0) { if (warp == 0) HMA_TL(5, it); if (warp == 3) HMA_TL(6, it); if (warp == 4) HMA_TL(7, it); if (warp == 7) HMA_TL(8, it); }
        }
      }
    } else {
      // ------------------------------------------------ stage B: (dP, P) -> dS; also writes dK / dV out
      auto store_dkdv = [&](int kt) {  // warps 8-11 store dK, warps 12-15 dV
        mbar_wait(smem_u32(&bar_kv), (uint32_t)(kt & 1));
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_x32((half == 0 ? tDK : tDV) + (uint32_t)(kt & 1) * kAccBuf + lane_addr, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_epi[kt & 1]));
        const int ki = kt * 128 + row;
        if (ki < n)
          store_head_row(p.dqkv + (size_t)(row0 + ki) * p.ld_dqkv + (half == 0 ? p.k_col : p.v_col) + head * 32, r);
      };
      int pending_kt = -1;
      int it = 0;
      for (int kt = 0; kt < ntile; ++kt) {
        const int nk = min(128, n - kt * 128);
        for (int qt = 0; qt < ntile; ++qt, ++it) {
          const int bsel = it & 1;
          if (it >= 2) mbar_wait(smem_u32(&bar_dsfree[bsel]), (uint32_t)(((it - 2) >> 1) & 1));  // dS[bsel] consumed by dK, dQ
          mbar_wait(smem_u32(&bar_dp), (uint32_t)(it & 1));
          tc_fence_after();
          mbar_wait(smem_u32(&bar_p[bsel]), (uint32_t)((it >> 1) & 1));  // P of this pair is in shared memory
          const int qi = qt * 128 + row;
          const float delta = qi < n ? s_delta[qi] : 0.f;
          const f32x2 sc2 = dup2(p.scale), nds2 = dup2(-delta * p.scale);
          const bool rows_live = quarter * 32 < min(128, n - qt * 128);
          const uint32_t sP = sP0 + (uint32_t)bsel * kPdsBuf + pan;
          const uint32_t sDS = sP + 2 * kBPanel;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t dp[32];
            tmem_ld_x32(tDP + lane_addr + (uint32_t)(half * 64 + c * 32), dp);
            tmem_ld_wait();
            if (c == 1) {
              tc_fence_before();
              mbar_arrive(smem_u32(&bar_dpfree));
            }
            if (rows_live && half * 64 + c * 32 < nk) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(c * 32 + q * 8));
                uint32_t pk[4];
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3])
                             : "r"(sP + off) : "memory");
                uint32_t dk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // dS = P * (scale * dP - scale * delta), two elements per instruction
                  const int e = q * 8 + 2 * j;
                  const f32x2 t = fma2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), sc2, nds2);
                  dk[j] = pack_bf16(mul2(pk2(bf16_lo(pk[j]), bf16_hi(pk[j])), t));
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(dk[0]), "r"(dk[1]), "r"(dk[2]), "r"(dk[3]) : "memory");
              }
            }
          }
          fence_proxy_async();
          mbar_arrive(smem_u32(&bar_ds[bsel]));
          if (lane == 0) { if (warp == 8) HMA_TL(12, it); if (warp == 9) HMA_TL(13, it); if (warp == 12) HMA_TL(14, it); if (warp == 15) HMA_TL(15, it); }
          // dK / dV of a finished key tile are written out one iteration late, after the dS of the next tile pair has
          // been handed to the tensor core, so the stores overlap its contractions
          if (pending_kt >= 0) {
            store_dkdv(pending_kt);
            pending_kt = -1;
          }
          if (qt == ntile - 1) pending_kt = kt;
        }
      }
      if (pending_kt >= 0) store_dkdv(pending_kt);
    }
    mbar_wait(smem_u32(&bar_final), 0);
    tc_fence_after();
    if (threadIdx.x == 0) HMA_TL(9, 0);
    const int dq_tile = warp >> 2;
    if (dq_tile < ntile) {  // warps 4*qt .. 4*qt+3 store the dQ rows of query tile qt
      uint32_t r[32];
      tmem_ld_x32(tDQ + (uint32_t)dq_tile * 32 + lane_addr, r);
      tmem_ld_wait();
      const int qi = dq_tile * 128 + row;
      if (qi < n) store_head_row(p.dqkv + (size_t)(row0 + qi) * p.ld_dqkv + p.q_col + head * 32, r);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HMA_TL(10, 0);
  if (warp == kComputeWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace hma

extern "C" int hma_attn_spatial_bwd(const void* qkv, long long ld_qkv, const void* out, long long ldo,
                                    const void* dout, long long ld_dout, const float* lse, int frames, int n,
                                    int heads, int q_col, int k_col, int v_col, float scale, void* dqkv,
                                    long long ld_dqkv, const float* delta, void* stream_) {
  using namespace hma;
  if (frames == 0) return 0;
  HMA_REQUIRE(n % 16 == 0 && n >= 16 && n <= kBMaxN, "attn_spatial_bwd: n=%d must be a multiple of 16 in [16,320]", n);
  HMA_REQUIRE(lse != nullptr, "attn_spatial_bwd: needs the forward log-sum-exp");
  AttnBwdParams p;
  p.n = n;
  p.box_rows = (n % 64 == 0) ? 64 : ((n % 32 == 0) ? 32 : 16);
  p.heads = heads;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<const __nv_bfloat16*>(out); p.ldo = ldo;
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.ld_dout = ld_dout;
  p.lse = lse;
  p.delta = delta;
  HMA_REQUIRE(delta != nullptr || out != nullptr, "attn_spatial_bwd: needs either delta or the forward output");
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  CUtensorMap tmQ, tmD;
  int rc = hma_host::make_tmap_bf16_2d_sw(&tmQ, qkv, (uint64_t)ld_qkv, (uint64_t)frames * n, (uint64_t)ld_qkv * 2, 32,
                                          (uint32_t)p.box_rows, 64);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d_sw(&tmD, dout, (uint64_t)ld_dout, (uint64_t)frames * n, (uint64_t)ld_dout * 2, 32,
                                      (uint32_t)p.box_rows, 64);
  if (rc) return rc;
  constexpr size_t smem = 1024 + 4 * kBTile + 8 * kBPanel;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_spatial_bwd_kernel, dim3(frames * heads), dim3(kBwdThreads), smem,
                                      static_cast<cudaStream_t>(stream_), tmQ, tmD, p));
  return 0;
}

HMA_DEFINE_TIMELINE_READER(hma_timeline_attn_spatial_bwd)
