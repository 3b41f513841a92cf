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
// Accumulators live in TMEM (S 128 + dP 128 + dQ 3x32 + dK 32 + dV 32 columns). Eight compute
// warps turn (S, dP) into (P, dS): warp w owns TMEM lane quarter w%4 and key-column half w/4.
// Software pipeline: the compute warps first copy their S / dP columns into registers and release
// TMEM, so the S / dP contractions of the NEXT tile pair run while they do the exp / dS math, and
// the P / dS tiles are double buffered in shared memory so they never wait for the dV/dK/dQ
// contractions of the previous pair either.
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
  __nv_bfloat16* dqkv;        // [tokens, ld_dqkv]
  long long ld_dqkv;
};

constexpr int kBRowB = 64;
constexpr int kBMaxN = 320;
constexpr int kBTile = kBMaxN * kBRowB;   // 20 KB per operand
constexpr int kBPanel = 128 * 128;        // [128 x 64] bf16 panel
constexpr int kComputeThreads = 256;

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

__global__ void __launch_bounds__(288, 1)
attn_spatial_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_sdp, bar_tfree, bar_pds[2], bar_mma[2], bar_kv, bar_epi, bar_final;
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
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_sdp), 1);
    mbar_init(smem_u32(&bar_tfree), kComputeThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_pds[i]), kComputeThreads);
      mbar_init(smem_u32(&bar_mma[i]), 1);
    }
    mbar_init(smem_u32(&bar_kv), 1);
    mbar_init(smem_u32(&bar_epi), kComputeThreads);
    mbar_init(smem_u32(&bar_final), 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (warp == 8) {
    // start the operand loads first: they overlap the per-row statistics below
    if (elect_one()) {
      const uint32_t bl = smem_u32(&bar_load);
      mbar_expect_tx(bl, (uint32_t)(4 * n * kBRowB));
      for (int r = 0; r < n; r += p.box_rows) {
        tma_load_2d(sQ + r * kBRowB, &tmQKV, bl, p.q_col + head * 32, row0 + r);
        tma_load_2d(sK + r * kBRowB, &tmQKV, bl, p.k_col + head * 32, row0 + r);
        tma_load_2d(sV + r * kBRowB, &tmQKV, bl, p.v_col + head * 32, row0 + r);
        tma_load_2d(sDO + r * kBRowB, &tmDO, bl, head * 32, row0 + r);
      }
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  // per-row statistics: log2-sum-exp and delta = dO . O
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    s_lse[r] = p.lse[((size_t)frame * p.heads + head) * n + r];
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDQ = tmem_base + 256, tDK = tmem_base + 352,
                 tDV = tmem_base + 384;

  if (warp == 8) {
    if (elect_one()) {
      mbar_wait(smem_u32(&bar_load), 0);
      tc_fence_after();

      auto issue_sdp = [&](int kt, int qt) {
        const int nk = min(128, n - kt * 128);
        const uint32_t idesc = umma_idesc_bf16(128, nk, 0, 0);
        const uint32_t q_addr = sQ + (uint32_t)qt * 128 * kBRowB;
        const uint32_t do_addr = sDO + (uint32_t)qt * 128 * kBRowB;
        const uint32_t k_addr = sK + (uint32_t)kt * 128 * kBRowB;
        const uint32_t v_addr = sV + (uint32_t)kt * 128 * kBRowB;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss(tS, bdesc_sw64(q_addr + k * 32), bdesc_sw64(k_addr + k * 32), idesc, (uint32_t)k);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss(tDP, bdesc_sw64(do_addr + k * 32), bdesc_sw64(v_addr + k * 32), idesc, (uint32_t)k);
      };

      const uint32_t idesc_t = umma_idesc_bf16(128, 32, 1, 1);   // dV, dK: both operands MN-major
      const uint32_t idesc_q = umma_idesc_bf16(128, 32, 0, 1);   // dQ: A K-major, B MN-major
      const int NI = ntile * ntile;
      issue_sdp(0, 0);
      umma_commit(smem_u32(&bar_sdp));
      for (int it = 0; it < NI; ++it) {
        const int kt = it / ntile, qt = it % ntile;
        const int nk = min(128, n - kt * 128);
        const int kq = min(128, n - qt * 128);
        const int bsel = it & 1;
        // S / dP of the next pair as soon as this pair's have been copied out of TMEM
        mbar_wait(smem_u32(&bar_tfree), (uint32_t)(it & 1));
        tc_fence_after();
        if (it + 1 < NI) {
          issue_sdp((it + 1) / ntile, (it + 1) % ntile);
          umma_commit(smem_u32(&bar_sdp));
        }
        mbar_wait(smem_u32(&bar_pds[bsel]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        if (qt == 0 && kt > 0) {
          mbar_wait(smem_u32(&bar_epi), (uint32_t)((kt - 1) & 1));
          tc_fence_after();
        }
        const uint32_t sP = sP0 + (uint32_t)bsel * kPdsBuf;
        const uint32_t sDS = sP + 2 * kBPanel;
        const uint32_t q_addr = sQ + (uint32_t)qt * 128 * kBRowB;
        const uint32_t do_addr = sDO + (uint32_t)qt * 128 * kBRowB;
        for (int kk = 0; kk < kq / 16; ++kk) {
          umma_ss(tDV, umma_desc_mnmajor(sP + kk * 2048, kBPanel), bdesc_sw64(do_addr + kk * 1024), idesc_t,
                  (uint32_t)((qt | kk) != 0));
          umma_ss(tDK, umma_desc_mnmajor(sDS + kk * 2048, kBPanel), bdesc_sw64(q_addr + kk * 1024), idesc_t,
                  (uint32_t)((qt | kk) != 0));
        }
        for (int kk = 0; kk < nk / 16; ++kk)
          umma_ss(tDQ + (uint32_t)qt * 32, umma_desc_kmajor(sDS + (uint32_t)(kk >> 2) * kBPanel + (uint32_t)(kk & 3) * 32),
                  bdesc_sw64(sK + (uint32_t)(kt * 128 + kk * 16) * kBRowB), idesc_q, (uint32_t)((kt | kk) != 0));
        umma_commit(smem_u32(&bar_mma[bsel]));
        if (qt == ntile - 1) umma_commit(smem_u32(&bar_kv));
        if (it == NI - 1) umma_commit(smem_u32(&bar_final));
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps 0..7
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int it = 0;
    for (int kt = 0; kt < ntile; ++kt) {
      const int nk = min(128, n - kt * 128);
      for (int qt = 0; qt < ntile; ++qt, ++it) {
        // (9 warps -> three share one SM sub-partition -> <= 168 registers: do every spin-wait BEFORE the
        //  128 S/dP registers become live so nothing spills inside a wait loop)
        const int bsel = it & 1;
        if (it >= 2) mbar_wait(smem_u32(&bar_mma[bsel]), (uint32_t)(((it - 2) >> 1) & 1));
        mbar_wait(smem_u32(&bar_sdp), (uint32_t)(it & 1));
        tc_fence_after();
        const int qi = qt * 128 + row;
        const float L = qi < n ? s_lse[qi] : 0.f;
        const float delta = qi < n ? s_delta[qi] : 0.f;
        // copy this warp's S / dP columns to registers and hand TMEM back to the tensor core
        uint32_t s[2][32], dp[2][32];
        const bool has0 = half * 64 < nk, has1 = half * 64 + 32 < nk;  // warp-uniform
        if (has0) {
          tmem_ld_x32(tS + lane_addr + half * 64, s[0]);
          tmem_ld_x32(tDP + lane_addr + half * 64, dp[0]);
        }
        if (has1) {
          tmem_ld_x32(tS + lane_addr + half * 64 + 32, s[1]);
          tmem_ld_x32(tDP + lane_addr + half * 64 + 32, dp[1]);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_tfree));
        const uint32_t sP = sP0 + (uint32_t)bsel * kPdsBuf;
        const uint32_t sDS = sP + 2 * kBPanel;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (cc == 0 ? !has0 : !has1) continue;
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = fast_ex2(fmaf(__uint_as_float(s[cc][j]), p.scale_log2, -L));
            const float p1 = fast_ex2(fmaf(__uint_as_float(s[cc][j + 1]), p.scale_log2, -L));
            const float d0 = p0 * (__uint_as_float(dp[cc][j]) - delta) * p.scale;
            const float d1 = p1 * (__uint_as_float(dp[cc][j + 1]) - delta) * p.scale;
            pk[j >> 1] = pack_bf16(p0, p1);
            dk[j >> 1] = pack_bf16(d0, d1);
          }
          const uint32_t pan = (uint32_t)half * kBPanel;
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
        mbar_arrive(smem_u32(&bar_pds[bsel]));
        if (qt == ntile - 1) {
          // dK / dV of this key tile are complete: warps 0-3 store dK, warps 4-7 store dV
          mbar_wait(smem_u32(&bar_kv), (uint32_t)(kt & 1));
          tc_fence_after();
          uint32_t r[32];
          tmem_ld_x32((half == 0 ? tDK : tDV) + lane_addr, r);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&bar_epi));
          const int ki = kt * 128 + row;
          if (ki < n)
            store_head_row(p.dqkv + (size_t)(row0 + ki) * p.ld_dqkv + (half == 0 ? p.k_col : p.v_col) + head * 32, r);
        }
      }
    }
    mbar_wait(smem_u32(&bar_final), 0);
    tc_fence_after();
    if (half == 0) {
      for (int qt = 0; qt < ntile; ++qt) {
        uint32_t r[32];
        tmem_ld_x32(tDQ + (uint32_t)qt * 32 + lane_addr, r);
        tmem_ld_wait();
        const int qi = qt * 128 + row;
        if (qi < n) store_head_row(p.dqkv + (size_t)(row0 + qi) * p.ld_dqkv + p.q_col + head * 32, r);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace hma

extern "C" int hma_attn_spatial_bwd(const void* qkv, long long ld_qkv, const void* out, long long ldo,
                                    const void* dout, long long ld_dout, const float* lse, int frames, int n,
                                    int heads, int q_col, int k_col, int v_col, float scale, void* dqkv,
                                    long long ld_dqkv, void* stream_) {
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
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  CUtensorMap tmQ, tmD;
  int rc = hma_host::make_tmap_bf16_2d_sw(&tmQ, qkv, (uint64_t)ld_qkv, (uint64_t)frames * n, (uint64_t)ld_qkv * 2, 32,
                                          (uint32_t)p.box_rows, 64);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d_sw(&tmD, dout, (uint64_t)ld_dout, (uint64_t)frames * n, (uint64_t)ld_dout * 2, 32,
                                      (uint32_t)p.box_rows, 64);
  if (rc) return rc;
  constexpr size_t smem = 1024 + 4 * kBTile + 8 * kBPanel;
  static bool attr_done = false;
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_spatial_bwd_kernel, dim3(frames * heads), dim3(288), smem,
                                      static_cast<cudaStream_t>(stream_), tmQ, tmD, p));
  return 0;
}
