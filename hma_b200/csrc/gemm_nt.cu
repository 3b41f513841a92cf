// D[M,N] = epilogue(A[M,K] . B[N,K]^T): the "x @ W^T" contraction used by every projection on
// the ST-transformer path (reference call sites: attention.py:141,154; st_transformer.py:24-27;
// st_mask_git.py:70-75,681-683 and their autograd transposes).
//
// sm_100a design: persistent CTAs, one 128 x BN output tile at a time.
//   warp 0   TMA producer  (cp.async.bulk.tensor, 128-byte swizzle, 64-wide K panels)
//   warp 1   UMMA issuer   (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue     (tcgen05.ld 32x32b: one output row per thread) with the fused
//                          bias / GELU / dGELU / fp32-residual variants
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1. When the whole B slice (BN x K) fits in shared memory it is loaded once
// per CTA and kept resident ("stationary"), so only A streams.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct GemmNtParams {
  int M, N, K;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const float* resid;
  long long ldr;
  const __nv_bfloat16* aux;
  long long ldaux;
  float alpha;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kAStage = kBM * kBK * 2;  // 16 KB
constexpr int kStages = 4;

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmNtParams& p, const uint32_t (&r)[32], int row,
                                               int n0) {
  // r: 32 consecutive fp32 accumulator columns [n0, n0+32) of output row `row`.
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (row >= p.M) return;

  if constexpr (EPI == HMA_EPI_BF16) {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  } else if constexpr (EPI == HMA_EPI_GELU_BF16 || EPI == HMA_EPI_SILU_BF16) {
    if (p.out2 != nullptr) {
      uint4* dz = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out2) + (size_t)row * p.ldo2 + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dz[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (EPI == HMA_EPI_GELU_BF16) ? gelu_erf(v[j]) : silu(v[j]);
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  } else if constexpr (EPI == HMA_EPI_DGELU_BF16 || EPI == HMA_EPI_DSILU_BF16) {
    const uint4* z = reinterpret_cast<const uint4*>(p.aux + (size_t)row * p.ldaux + n0);
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 zz = z[j];
      const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float d0 = (EPI == HMA_EPI_DGELU_BF16) ? dgelu_erf(bf16_lo(zw[q])) : dsilu(bf16_lo(zw[q]));
        const float d1 = (EPI == HMA_EPI_DGELU_BF16) ? dgelu_erf(bf16_hi(zw[q])) : dsilu(bf16_hi(zw[q]));
        const float g0 = v[8 * j + 2 * q] * d0;
        const float g1 = v[8 * j + 2 * q + 1] * d1;
        o[q] = pack_bf16(g0, g1);
      }
      dst[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else if constexpr (EPI == HMA_EPI_RESID_F32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)row * p.ldo + n0);
    if (p.resid != nullptr) {
      const float4* src = reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldr + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = src[j];
        dst[j] = make_float4(x.x + v[4 * j], x.y + v[4 * j + 1], x.z + v[4 * j + 2], x.w + v[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
}

template <int BN, int EPI, bool STAT>
__global__ void __launch_bounds__(256, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmNtParams p) {
  constexpr int kBStage = BN * kBK * 2;
  constexpr uint32_t kTmemCols = 2 * BN;
  constexpr uint32_t kIdesc = umma_idesc_bf16(kBM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages];
  __shared__ __align__(8) uint64_t bar_empty[kStages];
  __shared__ __align__(8) uint64_t bar_bfull;
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smemA = smem_base;
  const uint32_t smemB = smem_base + kStages * kAStage;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.K / kBK;
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int n_blk = blockIdx.x % n_tiles;
  const int m_start = blockIdx.x / n_tiles;
  const int m_step = gridDim.x / n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_bfull), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      if constexpr (STAT) {
        mbar_expect_tx(smem_u32(&bar_bfull), (uint32_t)(KB * kBStage));
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(smemB + kb * kBStage, &tmB, smem_u32(&bar_bfull), kb * kBK, n_blk * BN);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int m_blk = m_start; m_blk < m_tiles; m_blk += m_step) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, (uint32_t)(kAStage + (STAT ? 0 : kBStage)));
          tma_load_2d(smemA + stage * kAStage, &tmA, full, kb * kBK, m_blk * kBM);
          if constexpr (!STAT) tma_load_2d(smemB + stage * kBStage, &tmB, full, kb * kBK, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- UMMA issuer
    if (elect_one()) {
      if constexpr (STAT) {
        mbar_wait(smem_u32(&bar_bfull), 0);
        tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int m_blk = m_start; m_blk < m_tiles; m_blk += m_step, ++it) {
        const int as = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(smem_u32(&bar_tempty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smemA + stage * kAStage;
          const uint32_t b_addr = STAT ? (smemB + kb * kBStage) : (smemB + stage * kBStage);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            umma_ss(d_tmem, umma_desc_kmajor(a_addr + k * 32), umma_desc_kmajor(b_addr + k * 32), kIdesc,
                    (uint32_t)((kb | k) != 0));
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_tfull[as]));
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    int it = 0;
    for (int m_blk = m_start; m_blk < m_tiles; m_blk += m_step, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
      const int row = m_blk * kBM + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem_addr(tmem_base, (uint32_t)(ew * 32), (uint32_t)(as * BN + c * 32)), r);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, r, row, n_blk * BN + c * 32);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int EPI, bool STAT>
static int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmNtParams& p, cudaStream_t stream) {
  constexpr int kBStage = BN * kBK * 2;
  const int KB = p.K / kBK;
  const size_t smem = 1024 + (size_t)kStages * kAStage + (STAT ? (size_t)KB * kBStage : (size_t)kStages * kBStage);
  auto kern = gemm_nt_kernel<BN, EPI, STAT>;
  static bool attr_done = false;  // idempotent; racing threads set the same value
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_done = true;
  }
  HMA_REQUIRE(smem <= 227 * 1024 - 2048, "gemm_nt: shared memory request %zu too large", smem);
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  int per_n = hma_host::sm_count() / n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > m_tiles) per_n = m_tiles;
  const int grid = per_n * n_tiles;
  kern<<<grid, 256, smem, stream>>>(tmA, tmB, p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int EPI>
static int dispatch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmNtParams& p, int bn,
                       cudaStream_t stream) {
  const bool stat = (size_t)p.K * bn * 2 <= 128 * 1024;
  if (bn == 256) {
    return stat ? launch_nt<256, EPI, true>(tmA, tmB, p, stream) : launch_nt<256, EPI, false>(tmA, tmB, p, stream);
  }
  return stat ? launch_nt<128, EPI, true>(tmA, tmB, p, stream) : launch_nt<128, EPI, false>(tmA, tmB, p, stream);
}

}  // namespace hma

extern "C" int hma_gemm_nt(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                           int epi, void* out, long long ldo, void* out2, long long ldo2, const float* bias,
                           const float* resid, long long ldr, const void* aux, long long ldaux, float alpha,
                           void* stream_) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M == 0) return 0;
  HMA_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_nt: bad shape M=%d N=%d K=%d", M, N, K);
  HMA_REQUIRE(K % kBK == 0, "gemm_nt: K=%d must be a multiple of 64", K);
  HMA_REQUIRE(N % 128 == 0, "gemm_nt: N=%d must be a multiple of 128", N);
  HMA_REQUIRE(out != nullptr, "gemm_nt: out is null");
  const int bn = (N % 256 == 0 && N >= 512) ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc = hma_host::make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBK, kBM);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBK, bn);
  if (rc) return rc;
  GemmNtParams p;
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldo = ldo; p.out2 = out2; p.ldo2 = ldo2;
  p.bias = bias; p.resid = resid; p.ldr = ldr;
  p.aux = static_cast<const __nv_bfloat16*>(aux); p.ldaux = ldaux;
  p.alpha = alpha;
  switch (epi) {
    case HMA_EPI_BF16: return dispatch_nt<HMA_EPI_BF16>(tmA, tmB, p, bn, stream);
    case HMA_EPI_GELU_BF16: return dispatch_nt<HMA_EPI_GELU_BF16>(tmA, tmB, p, bn, stream);
    case HMA_EPI_DGELU_BF16:
      HMA_REQUIRE(aux != nullptr, "gemm_nt: dGELU epilogue needs the saved pre-activation");
      return dispatch_nt<HMA_EPI_DGELU_BF16>(tmA, tmB, p, bn, stream);
    case HMA_EPI_RESID_F32: return dispatch_nt<HMA_EPI_RESID_F32>(tmA, tmB, p, bn, stream);
    case HMA_EPI_SILU_BF16: return dispatch_nt<HMA_EPI_SILU_BF16>(tmA, tmB, p, bn, stream);
    case HMA_EPI_DSILU_BF16:
      HMA_REQUIRE(aux != nullptr, "gemm_nt: dSiLU epilogue needs the saved pre-activation");
      return dispatch_nt<HMA_EPI_DSILU_BF16>(tmA, tmB, p, bn, stream);
    default: break;
  }
  HMA_REQUIRE(false, "gemm_nt: unknown epilogue %d", epi);
}
