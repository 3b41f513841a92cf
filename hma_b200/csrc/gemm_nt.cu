// D[M,N] = epilogue(A[M,K] . B[N,K]^T): the "x @ W^T" contraction used by every projection on
// the ST-transformer path (reference call sites: attention.py:141,154; st_transformer.py:24-27;
// st_mask_git.py:70-75,681-683 and their autograd transposes).
//
// sm_100a design: persistent CTAs, one 128 x BN output tile at a time.
//   warp 0     TMA producer  (cp.async.bulk.tensor, 128-byte swizzle, 64-wide K panels)
//   warp 1     UMMA issuer   (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM)
//   warp 2     TMEM allocator
//   warps 4-11 epilogue      (tcgen05.ld 32x32b: one accumulator row per thread; two warps per TMEM
//                            lane quarter, alternating 32-column chunks)
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1. When the whole B slice (BN x K) fits in shared memory it is loaded once
// per CTA and kept resident ("stationary"), so only A streams.
//
// Epilogue memory access: a thread owns a ROW of the accumulator, which is the worst possible
// shape for global memory (32 lanes -> 32 different rows). Every chunk is therefore transposed
// through a per-warp padded shared-memory buffer, and residual / saved-activation reads and all
// stores are issued in the transposed domain, where a warp instruction touches whole 64/128-byte
// row segments.
//
// Tile width BN: 256 for N >= 512, 128 otherwise, 64 when the launch has so few row tiles that wider tiles would leave most
// SMs idle (sampler / small-batch decode shapes). Element-wise epilogue arithmetic works on packed fp32 pairs (FFMA2).
// LN = true (hma_gemm_nt_ln): the residual epilogue also emits the next stage's LayerNorm; the two 128-column CTAs of a row
// tile form a thread-block cluster and exchange their halves of the row sums through distributed shared memory.
#include <cstdio>
#include <cstdlib>
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
  float* colsum;  // d-activation epilogues: colsum[N] += column sums of the output (the bias gradient of the Linear below)
  float* rowdot;  // EPI_BF16: rowdot[row, N/32] = sum over each 32-column chunk of bf16(out) * aux (attention backward's delta)
  // 3x3 convolution as ONE contraction (hma_conv3x3_nhwc): A is a zero-bordered NHWC image [rows, Cin] and K = 9 * Cin; the
  // k-blocks of tap (ky, kx) are the SAME columns of A read (ky - 1) * conv_wp + (kx - 1) rows further down — a TMA
  // coordinate, not an im2col copy. conv_kb_per_tap = Cin / 64; 0 = ordinary GEMM.
  int conv_kb_per_tap;
  int conv_wp;
  int stages;  // depth of the A (and streamed B) ring, 2..kMaxStages: as deep as shared memory allows (bytes in flight per SM)
  // LayerNorm of the output rows, emitted by the residual epilogue itself (LN kernels: N == 256 as two 128-column CTAs in a
  // thread-block cluster, which exchange their halves of the row sums through distributed shared memory): ln_out = bf16(LN(out)) in the flavour the NEXT stage wants — mode 1: affine (gamma, beta); mode 2: no affine,
  // modulated per group of ln_rpg rows by mod[group] = shift[256] | scale[256] (st_mask_git.py:66-76). ln_stats (optional)
  // receives (mean, rstd) per row for the backward.
  int ln_mode;
  const float* ln_gamma;
  const float* ln_beta;
  const float* ln_mod;
  int ln_rpg;
  float ln_eps;
  __nv_bfloat16* ln_out;
  long long ln_ldo;
  float* ln_stats;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kAStage = kBM * kBK * 2;      // 16 KB
// Epilogue warps: 8 (two per TMEM lane quarter), or 16 for the GELU epilogue, whose arithmetic (erf, two bf16
// outputs per accumulator) is issue- and latency-bound with only two warps per scheduler.
template <int EPI> struct EpiWarps { static constexpr int value = (EPI == HMA_EPI_GELU_BF16 || EPI == HMA_EPI_DGELU_BF16) ? 16 : 8; };
// Per-warp staging buffer: 32 padded fp32 rows when the epilogue transposes fp32 (residual / d-activation), else bf16 rows.
template <int EPI> struct StageBytes {
  static constexpr int value = (EPI == HMA_EPI_RESID_F32 || EPI == HMA_EPI_DGELU_BF16 || EPI == HMA_EPI_DSILU_BF16) ? 32 * 144 : 32 * 80;
};
constexpr int kStageF32Row = 144;           // 32 fp32 + 16 B pad: conflict-free 16-byte row writes
constexpr int kStageBf16Row = 80;           // 32 bf16 + 16 B pad
constexpr int kSmemLimit = 227 * 1024 - 1024;
constexpr int kMaxStages = 8;
constexpr int kLnStash = 8 * 8192;          // LN epilogue: 8 warps x (2 chunks x 32 x 32 fp32)
constexpr int kLnStatic = 10 * 1024;        // LN epilogue: row-sum exchange buffers (static shared memory)

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Phi(z) = 0.5 (1 + erf(z / sqrt 2)) as a logistic of an odd polynomial: Phi(z) ~= 1 / (1 + 2^(z (a1 + a3 z^2 + a5 z^4))),
// coefficients fitted to logit(Phi) on |z| <= 5.5 (max |error| 3.7e-5 on Phi, 3.0e-5 on z Phi(z): below the bf16
// resolution of every consumer). 8 instructions (2 MUFU) against 14 for the Abramowitz-Stegun erf it replaces; the
// GELU epilogue is instruction-bound (ncu: 22.8 instructions per output element, 56 % issue utilisation).
// z^2 is clamped for the polynomial so that the quartic term cannot turn the logit around for |z| > 11.
// 1 / (1 + 2^u) = 0.5 - 0.5 tanh(u ln2 / 2): ONE MUFU op (tanh.approx, relative error 2^-11 -> |error| <= 2.5e-4 on Phi, an
// order of magnitude under the bf16 resolution of every consumer) instead of ex2 + rcp; the GELU / dGELU epilogues are
// MUFU- and issue-bound (3 -> 2 MUFU ops per element in the backward, 2 -> 1 in the forward).
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_cdf(float z, float zz) {
  const float z2 = fminf(zz, 64.0f);
  float t = fmaf(0.00099209175f * 0.34657359f, z2, -0.10660493f * 0.34657359f);
  t = fmaf(t, z2, -2.3013592f * 0.34657359f);
  return fmaf(-0.5f, fast_tanh(t * z), 0.5f);
}
__device__ __forceinline__ void gelu_parts(float z, float& cdf, float& pdf) {
  const float zz = z * z;
  cdf = gelu_cdf(z, zz);
  pdf = 0.3989422804014327f * fast_ex2(-0.72134752044448170f * zz);
}
// the same on a pair of values: 6 packed fp32 instructions, 2 FMNMX and 2 MUFU for two GELUs
__device__ __forceinline__ f32x2 gelu_cdf2(f32x2 z, f32x2 zz) {
  float a, b;
  upk2(zz, a, b);
  const f32x2 z2 = pk2(fminf(a, 64.0f), fminf(b, 64.0f));
  f32x2 t = fma2(dup2(0.00099209175f * 0.34657359f), z2, dup2(-0.10660493f * 0.34657359f));
  t = fma2(t, z2, dup2(-2.3013592f * 0.34657359f));
  upk2(mul2(t, z), a, b);
  return fma2(dup2(-0.5f), pk2(fast_tanh(a), fast_tanh(b)), dup2(0.5f));
}
__device__ __forceinline__ f32x2 gelu2(f32x2 z) { return mul2(z, gelu_cdf2(z, mul2(z, z))); }
// d gelu / dz = Phi(z) + z phi(z)
__device__ __forceinline__ f32x2 dgelu2(f32x2 z) {
  const f32x2 zz = mul2(z, z);
  const f32x2 cdf = gelu_cdf2(z, zz);
  float a, b;
  upk2(mul2(zz, dup2(-0.72134752044448170f)), a, b);
  return fma2(mul2(z, dup2(0.3989422804014327f)), pk2(fast_ex2(a), fast_ex2(b)), cdf);
}
template <int EPI>
__device__ __forceinline__ float act_fwd(float z) {
  if constexpr (EPI == HMA_EPI_GELU_BF16) {
    return z * gelu_cdf(z, z * z);
  } else {
    return silu(z);
  }
}
template <int EPI>
__device__ __forceinline__ float act_bwd(float z) {
  if constexpr (EPI == HMA_EPI_DGELU_BF16) {
    float cdf, pdf;
    gelu_parts(z, cdf, pdf);
    return fmaf(z, pdf, cdf);
  } else {
    return dsilu(z);
  }
}

// Stage 32 fp32 values of this lane's row, then hand each lane 4 consecutive columns of row
// (4*i + lane/8), i = 0..7, via `f(row_in_chunk, col_in_chunk, float4)`.
template <class F>
__device__ __forceinline__ void transpose_f32(uint32_t stage, int lane, const float (&v)[32], F&& f) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    sts_v4(stage + lane * kStageF32Row + q * 16, __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
           __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
  __syncwarp();
  // all eight shared-memory reads are issued before the first consumer: the consumers store to global memory, and the
  // compiler keeps those stores and the (volatile) shared loads in program order — interleaved, every row would pay the
  // shared-memory latency again
  uint4 u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = lds_v4(stage + (4 * i + (lane >> 3)) * kStageF32Row + (lane & 7) * 16);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    f(4 * i + (lane >> 3), (lane & 7) * 4,
      make_float4(__uint_as_float(u[i].x), __uint_as_float(u[i].y), __uint_as_float(u[i].z), __uint_as_float(u[i].w)));
  __syncwarp();
}

// Stage 32 bf16 values of this lane's row and store the chunk to `dst` (row-major bf16, leading
// dimension ld) with 64-byte row segments: lane handles row (8*i + lane/4), 16-byte piece lane%4.
__device__ __forceinline__ void store_chunk_bf16(uint32_t stage, int lane, const float (&v)[32], __nv_bfloat16* dst,
                                                 long long ld, int row0, int n0, int M) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    sts_v4(stage + lane * kStageBf16Row + q * 16, pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
           pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
  __syncwarp();
  uint4 u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = lds_v4(stage + (8 * i + (lane >> 2)) * kStageBf16Row + (lane & 3) * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = 8 * i + (lane >> 2);
    if (row0 + rr < M) *reinterpret_cast<uint4*>(dst + (size_t)(row0 + rr) * ld + n0 + (lane & 3) * 8) = u[i];
  }
  __syncwarp();
}

// Positions a lane owns in the transposed domain of a 32x32 chunk: row 4*i + lane/8, columns (lane%8)*4..+3.
template <int EPI>
__device__ __forceinline__ void prefetch_chunk(const GemmNtParams& p, int row0, int n0, int lane, float4 (&pf)[8]) {
  // RESID: the fp32 residual; DGELU/DSILU: the saved bf16 pre-activation (in the low 8 bytes)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + 4 * i + (lane >> 3);
    const int cc = (lane & 7) * 4;
    pf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < p.M) {
      if constexpr (EPI == HMA_EPI_RESID_F32) {
        if (p.resid != nullptr) pf[i] = *reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldr + n0 + cc);
      } else {
        const uint2 z = *reinterpret_cast<const uint2*>(p.aux + (size_t)row * p.ldaux + n0 + cc);
        pf[i].x = __uint_as_float(z.x);
        pf[i].y = __uint_as_float(z.y);
      }
    }
  }
}

// rowdot operand: the 32 bf16 of aux[row0 + lane, n0 .. n0+32) (row-owner domain), fetched one chunk ahead like pf
__device__ __forceinline__ void prefetch_rowdot(const GemmNtParams& p, int row0, int n0, int lane, uint4 (&po)[4]) {
  const int row = row0 + lane;
#pragma unroll
  for (int q = 0; q < 4; ++q) po[q] = make_uint4(0u, 0u, 0u, 0u);
  if (row < p.M) {
    const uint4* o4 = reinterpret_cast<const uint4*>(p.aux + (size_t)row * p.ldaux + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) po[q] = __ldg(o4 + q);
  }
}

template <int EPI, bool LN>
__device__ __forceinline__ void epilogue_chunk(const GemmNtParams& p, const uint32_t (&r)[32], int row0, int n0,
                                               int lane, uint32_t stage, const float4 (&pf)[8], float4& csum,
                                               const uint4 (&po)[4], f32x2 (&s1)[8], f32x2 (&s2)[8], uint32_t stash) {
  // r: 32 consecutive fp32 accumulator columns [n0, n0+32) of output row (row0 + lane).
  float v[32];
  {
    const f32x2 al = dup2(p.alpha);
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        upk2(fma2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), al, pk2(b.x, b.y)), v[j], v[j + 1]);
        upk2(fma2(pk2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), al, pk2(b.z, b.w)), v[j + 2], v[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) upk2(mul2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), al), v[j], v[j + 1]);
    }
  }
  if constexpr (EPI == HMA_EPI_BF16) {
    if (p.rowdot != nullptr) {
      // delta[row, head] = sum_c dO[row, c] O[row, c] over the 32 channels of a head (= this chunk), taken on the bf16
      // values the attention backward will read. A lane owns the row here; O's 64 bytes of the row are two full sectors.
      const int row = row0 + lane;
      if (row < p.M) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t ow[4] = {po[q].x, po[q].y, po[q].z, po[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t g = pack_bf16(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
            acc += bf16_lo(g) * bf16_lo(ow[j]) + bf16_hi(g) * bf16_hi(ow[j]);
          }
        }
        p.rowdot[(size_t)row * (p.N >> 5) + (n0 >> 5)] = acc;
      }
    }
    store_chunk_bf16(stage, lane, v, static_cast<__nv_bfloat16*>(p.out), p.ldo, row0, n0, p.M);
  } else if constexpr (EPI == HMA_EPI_GELU_BF16 || EPI == HMA_EPI_SILU_BF16) {
    if (p.out2 != nullptr) store_chunk_bf16(stage, lane, v, static_cast<__nv_bfloat16*>(p.out2), p.ldo2, row0, n0, p.M);
    if constexpr (EPI == HMA_EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) upk2(gelu2(pk2(v[j], v[j + 1])), v[j], v[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = act_fwd<EPI>(v[j]);
    }
    store_chunk_bf16(stage, lane, v, static_cast<__nv_bfloat16*>(p.out), p.ldo, row0, n0, p.M);
  } else if constexpr (EPI == HMA_EPI_DGELU_BF16 || EPI == HMA_EPI_DSILU_BF16) {
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
    transpose_f32(stage, lane, v, [&](int rr, int cc, float4 a) {
      const int row = row0 + rr;
      if (row < p.M) {
        const float4 zf = pf[rr >> 2];
        const uint2 z = make_uint2(__float_as_uint(zf.x), __float_as_uint(zf.y));
        float g0, g1, g2, g3;
        if constexpr (EPI == HMA_EPI_DGELU_BF16) {
          upk2(mul2(pk2(a.x, a.y), dgelu2(pk2(bf16_lo(z.x), bf16_hi(z.x)))), g0, g1);
          upk2(mul2(pk2(a.z, a.w), dgelu2(pk2(bf16_lo(z.y), bf16_hi(z.y)))), g2, g3);
        } else {
          g0 = a.x * act_bwd<EPI>(bf16_lo(z.x)); g1 = a.y * act_bwd<EPI>(bf16_hi(z.x));
          g2 = a.z * act_bwd<EPI>(bf16_lo(z.y)); g3 = a.w * act_bwd<EPI>(bf16_hi(z.y));
        }
        *reinterpret_cast<uint2*>(out + (size_t)row * p.ldo + n0 + cc) = make_uint2(pack_bf16(g0, g1), pack_bf16(g2, g3));
        csum.x += g0; csum.y += g1; csum.z += g2; csum.w += g3;  // this lane's 4 columns of the chunk, over its 8 rows
      }
    });
  } else if constexpr (EPI == HMA_EPI_RESID_F32) {
    float* out = static_cast<float*>(p.out);
    __nv_bfloat16* out2 = static_cast<__nv_bfloat16*>(p.out2);  // optional bf16 copy: the next stage's GEMM operand
    transpose_f32(stage, lane, v, [&](int rr, int cc, float4 a) {
      const int row = row0 + rr;
      if (row < p.M) {
        const float4 x = pf[rr >> 2];
        upk2(add2(pk2(a.x, a.y), pk2(x.x, x.y)), a.x, a.y);
        upk2(add2(pk2(a.z, a.w), pk2(x.z, x.w)), a.z, a.w);
        *reinterpret_cast<float4*>(out + (size_t)row * p.ldo + n0 + cc) = a;
        if constexpr (LN) {  // this lane's share of the row sums (rows 4 i + lane / 8, i = rr / 4); x itself stays on chip
          sts_v4(stash + (uint32_t)(((rr >> 2) * 32 + lane) * 16), __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z),
                 __float_as_uint(a.w));
          const f32x2 lo = pk2(a.x, a.y), hi = pk2(a.z, a.w);
          s1[rr >> 2] = add2(s1[rr >> 2], add2(lo, hi));
          s2[rr >> 2] = fma2(lo, lo, fma2(hi, hi, s2[rr >> 2]));
        } else if (out2 != nullptr) {
          const uint2 b = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
          *reinterpret_cast<uint2*>(out2 + (size_t)row * p.ldo2 + n0 + cc) = b;
          csum.x += bf16_lo(b.x); csum.y += bf16_hi(b.x); csum.z += bf16_lo(b.y); csum.w += bf16_hi(b.y);
        }
      }
    });
  }
}

template <int BN, int EPI, bool STAT, bool LN = false>
__global__ void __launch_bounds__(128 + 32 * EpiWarps<EPI>::value, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmNtParams p) {
  constexpr int kBStage = BN * kBK * 2;
  constexpr int kEpiWarps = EpiWarps<EPI>::value;
  constexpr int kStageBytes = StageBytes<EPI>::value;
  constexpr int kColGroups = kEpiWarps / 4;  // warps sharing a TMEM lane quarter split the 32-column chunks
  constexpr uint32_t kTmemCols = 2 * BN;
  constexpr uint32_t kIdesc = umma_idesc_bf16(kBM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_bfull;
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  // LN: a row's (sum, sum of squares) arrive in four parts — two warps (column groups) in each of the two CTAs of the cluster:
  // ln_xchg[tile parity][part = 2 * cta rank + column group][128 rows], every part written by its owner into BOTH CTAs
  // (its own and, through distributed shared memory, the peer's) by st.async, counted in bytes on ln_xbar[parity][lane quarter]
  __shared__ __align__(16) float ln_xchg[LN ? 2 * 4 * 128 * 2 : 4];
  __shared__ __align__(8) uint64_t ln_xbar[2][4];

  const int KB = p.K / kBK;
  const int kStages = p.stages;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smemA = smem_base;
  const uint32_t smemB = smem_base + kStages * kAStage;
  const uint32_t smemStage = smemB + (STAT ? KB : kStages) * kBStage;
  const uint32_t smemStash = smemStage + (uint32_t)(kEpiWarps * kStageBytes);  // LN: fp32 x of the tile, 8 KB per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
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
      mbar_init(smem_u32(&bar_tempty[s]), kEpiWarps * 32);
      if constexpr (LN) {
        for (int q = 0; q < 4; ++q) mbar_init(smem_u32(&ln_xbar[s][q]), 1);  // one expect_tx per tile; the parts count in bytes
      }
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
  if constexpr (LN) cluster_sync();  // the peer's barriers exist before anything is sent to them
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();               // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

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
          int a_col = kb * kBK, a_row = m_blk * kBM;
          if (p.conv_kb_per_tap > 0) {
            const int tap = kb / p.conv_kb_per_tap;
            a_col = (kb - tap * p.conv_kb_per_tap) * kBK;
            a_row += (tap / 3 - 1) * p.conv_wp + (tap % 3 - 1);  // rows before the image / past its end read as zero
          }
          tma_load_2d(smemA + stage * kAStage, &tmA, full, a_col, a_row);
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
    const int ew = warp & 3;         // TMEM lane quarter this warp may read
    const int eh = (warp - 4) >> 2;  // which 32-column chunks (c % kColGroups == eh)
    const uint32_t stage_buf = smemStage + (uint32_t)(warp - 4) * kStageBytes;
    // Global operands of the epilogue (fp32 residual / saved pre-activation) do not depend on the
    // accumulator: they are fetched one chunk ahead, the first chunk of a tile BEFORE waiting for the
    // MMAs, so their HBM latency hides behind the tensor-core work.
    constexpr bool kPrefetch = (EPI == HMA_EPI_RESID_F32 || EPI == HMA_EPI_DGELU_BF16 || EPI == HMA_EPI_DSILU_BF16);
    constexpr int kChunks = BN / 32 / kColGroups;  // chunks per warp and tile
    // bias gradient of the layer below (d-activation epilogues): a CTA keeps one n-block, so each lane accumulates
    // the column sums of its 4 columns of every chunk over all of the CTA's row tiles in registers
    float4 csum[kChunks];
#pragma unroll
    for (int j = 0; j < kChunks; ++j) csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    static_assert(!LN || (EPI == HMA_EPI_RESID_F32 && BN == 128 && kColGroups == 2 && kChunks == 2),
                  "LN epilogue: half rows per CTA, two warps per quarter");
    const uint32_t stash = smemStash + (uint32_t)(warp - 4) * 8192u;
    const int rq = lane >> 3;
    f32x2 s1[8], s2[8];  // LN: this lane's row sums (as pairs) of the tile in flight
    float4 ln_g[kChunks], ln_b[kChunks];  // LN: scale / shift of this lane's columns for the pending tile (fetched early)
    bool ln_one_group = true;
    int pend_row0 = -1, pend_it = 0;
    // Second half of the LayerNorm of a tile, deferred until the NEXT tile's residual prefetch is in flight so that the
    // exchange of the row sums hides behind it: full-row statistics (lane r: row r of the warp's 32), then y = LN(x) from the
    // stashed x in the transposed layout (lane: rows 4 i + lane / 8, four columns).
    auto ln_finish = [&]() {
      const int b = pend_it & 1;
      mbar_wait(smem_u32(&ln_xbar[b][ew]), (uint32_t)(pend_it >> 1) & 1u);
      float rs, nm;
      {
        const float* src = ln_xchg + ((b * 4) * 128 + ew * 32 + lane) * 2;
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int part = 0; part < 4; ++part) {  // fixed order: every warp that holds a piece of the row gets the same bits
          const float2 o2 = *reinterpret_cast<const float2*>(src + part * 256);
          t1 += o2.x;
          t2 += o2.y;
        }
        const float mean = t1 * (1.0f / 256.0f);
        rs = rsqrtf(fmaxf(t2 * (1.0f / 256.0f) - mean * mean, 0.f) + p.ln_eps);
        nm = -mean * rs;
        const int row = pend_row0 + lane;
        if (p.ln_stats != nullptr && n_blk == 0 && eh == 0 && row < p.M)
          *reinterpret_cast<float2*>(p.ln_stats + (size_t)row * 2) = make_float2(mean, rs);
      }
      const int cc = (lane & 7) * 4;
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {
        const int n0 = n_blk * BN + (eh + kColGroups * j) * 32 + cc;
        uint4 u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = lds_v4(stash + (uint32_t)(((j * 8 + i) * 32 + lane) * 16));
        float4 g = ln_g[j], bt = ln_b[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const f32x2 rs2 = dup2(__shfl_sync(0xffffffffu, rs, 4 * i + rq)), nm2 = dup2(__shfl_sync(0xffffffffu, nm, 4 * i + rq));
          const int row = pend_row0 + 4 * i + rq;
          if (row >= p.M) continue;
          if (!ln_one_group) {  // a warp's 32 rows straddle two modulation groups: per-row shift / scale
            const float* m = p.ln_mod + (size_t)(row / p.ln_rpg) * 512;
            bt = __ldg(reinterpret_cast<const float4*>(m + n0));
            g = __ldg(reinterpret_cast<const float4*>(m + 256 + n0));
            g.x += 1.f; g.y += 1.f; g.z += 1.f; g.w += 1.f;
          }
          const f32x2 y01 = fma2(fma2(pk2(__uint_as_float(u[i].x), __uint_as_float(u[i].y)), rs2, nm2), pk2(g.x, g.y), pk2(bt.x, bt.y));
          const f32x2 y23 = fma2(fma2(pk2(__uint_as_float(u[i].z), __uint_as_float(u[i].w)), rs2, nm2), pk2(g.z, g.w), pk2(bt.z, bt.w));
          *reinterpret_cast<uint2*>(p.ln_out + (size_t)row * p.ln_ldo + n0) = make_uint2(pack_bf16(y01), pack_bf16(y23));
        }
      }
    };
    int it = 0;
    for (int m_blk = m_start; m_blk < m_tiles; m_blk += m_step, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int row0 = m_blk * kBM + ew * 32;
      float4 pf[2][8];
      uint4 po[2][4];
      if constexpr (kPrefetch) prefetch_chunk<EPI>(p, row0, n_blk * BN + eh * 32, lane, pf[0]);
      if constexpr (EPI == HMA_EPI_BF16) {
        if (p.rowdot != nullptr) prefetch_rowdot(p, row0, n_blk * BN + eh * 32, lane, po[0]);
      }
      if constexpr (LN) {
        if (pend_row0 >= 0) ln_finish();
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1[i] = 0ull; s2[i] = 0ull; }
      }
      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {
        const int c = eh + kColGroups * j;
        if constexpr (kPrefetch) {
          if (j + 1 < kChunks) prefetch_chunk<EPI>(p, row0, n_blk * BN + (c + kColGroups) * 32, lane, pf[(j + 1) & 1]);
        }
        if constexpr (EPI == HMA_EPI_BF16) {
          if (p.rowdot != nullptr && j + 1 < kChunks) prefetch_rowdot(p, row0, n_blk * BN + (c + kColGroups) * 32, lane, po[(j + 1) & 1]);
        }
        uint32_t r[32];
        tmem_ld_x32(tmem_addr(tmem_base, (uint32_t)(ew * 32), (uint32_t)(as * BN + c * 32)), r);
        tmem_ld_wait();
        epilogue_chunk<EPI, LN>(p, r, row0, n_blk * BN + c * 32, lane, stage_buf, pf[j & 1], csum[j], po[j & 1], s1, s2,
                                stash + (uint32_t)j * 4096u);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
      if constexpr (LN) {
        // ---- first half of the LayerNorm of the 128 x 128 half-rows just written: this warp's part of the row sums goes to
        // both CTAs of the cluster; nobody waits here. The 8 lanes that share a row are summed through the warp's staging
        // buffer (value k of lane l at word 33 k + l: conflict-free both ways), which leaves row r's sums in lane r.
        {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float lo, hi;
            upk2(s1[i], lo, hi);
            const float a1 = lo + hi;
            upk2(s2[i], lo, hi);
            const float a2 = lo + hi;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage_buf + (uint32_t)((33 * i + lane) * 4)), "f"(a1) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage_buf + (uint32_t)((33 * (8 + i) + lane) * 4)), "f"(a2) : "memory");
          }
          __syncwarp();
          float t1 = 0.f, t2 = 0.f;  // lane r = row r = 4 i + rq': its parts sit in lanes 8 rq' .. 8 rq' + 7, value index i
          const uint32_t src = stage_buf + (uint32_t)((33 * (lane >> 2) + 8 * (lane & 3)) * 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float x1, x2;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(src + (uint32_t)(k * 4)) : "memory");
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x2) : "r"(src + (uint32_t)((33 * 8 + k) * 4)) : "memory");
            t1 += x1;
            t2 += x2;
          }
          __syncwarp();
          const int b = it & 1;
          const uint32_t slot = smem_u32(ln_xchg + ((b * 4 + (int)cluster_ctarank() * 2 + eh) * 128 + ew * 32 + lane) * 2);
          const uint32_t bar = smem_u32(&ln_xbar[b][ew]);
          if (eh == 0 && lane == 0) mbar_expect_tx(bar, 4 * 32 * 8);  // this CTA's barrier: 4 parts x 32 rows x 8 bytes
#pragma unroll
          for (int r = 0; r < 2; ++r) st_async_f32x2(mapa_shared(slot, (uint32_t)r), t1, t2, mapa_shared(bar, (uint32_t)r));
        }
        // scale / shift of this lane's columns for the second half, fetched now so that their latency is long gone by then
        {
          const int cc = (lane & 7) * 4;
          const int g_first = p.ln_mode == 2 ? row0 / p.ln_rpg : 0;
          ln_one_group = p.ln_mode != 2 || g_first == min(row0 + 31, p.M - 1) / p.ln_rpg;
#pragma unroll
          for (int j = 0; j < kChunks; ++j) {
            const int n0 = n_blk * BN + (eh + kColGroups * j) * 32 + cc;
            ln_g[j] = make_float4(1.f, 1.f, 1.f, 1.f);
            ln_b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.ln_mode == 1) {
              ln_g[j] = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + n0));
              ln_b[j] = __ldg(reinterpret_cast<const float4*>(p.ln_beta + n0));
            } else if (ln_one_group && row0 < p.M) {
              const float* m = p.ln_mod + (size_t)g_first * 512;
              ln_b[j] = __ldg(reinterpret_cast<const float4*>(m + n0));
              ln_g[j] = __ldg(reinterpret_cast<const float4*>(m + 256 + n0));
              ln_g[j].x += 1.f; ln_g[j].y += 1.f; ln_g[j].z += 1.f; ln_g[j].w += 1.f;
            }
          }
        }
        pend_row0 = row0;
        pend_it = it;
      }
    }
    if constexpr (LN) {
      if (pend_row0 >= 0) ln_finish();
    }
    if constexpr (EPI == HMA_EPI_DGELU_BF16 || EPI == HMA_EPI_DSILU_BF16 || EPI == HMA_EPI_RESID_F32) {
      if (p.colsum != nullptr) {
#pragma unroll
        for (int j = 0; j < kChunks; ++j) {
          float4 t = csum[j];  // lanes with equal lane % 8 hold the same columns for different rows
#pragma unroll
          for (int o = 8; o < 32; o <<= 1) {
            t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
            t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
            t.z += __shfl_xor_sync(0xffffffffu, t.z, o);
            t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
          }
          if (lane < 8) {
            float* dst = p.colsum + n_blk * BN + (eh + kColGroups * j) * 32 + lane * 4;
            atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if constexpr (LN) cluster_sync();  // neither CTA leaves while the other may still address its shared memory
}

template <int BN, int EPI>
static size_t smem_need(int KB, bool stat, int stages) {
  constexpr int kBStage = BN * kBK * 2;
  return 1024 + (size_t)stages * kAStage + (size_t)(stat ? KB : stages) * kBStage +
         (size_t)EpiWarps<EPI>::value * StageBytes<EPI>::value;
}
// Deepest ring that fits: the kernels are HBM-bound and a CTA's bytes in flight are what it can ask of the memory system
// (3 x 16 KB was ~Little's-law minimum at an unloaded latency; under load the latency doubles).
template <int BN, int EPI>
static int pick_stages(int KB, bool stat) {
  // measured (tools/kbench.py, config-2 shapes): the deep ring helps where A dominates the traffic (bf16 epilogues 18.6 ->
  // 16.7 us, the K = 1024 residual GEMM 50 -> 45 us) and HURTS the stationary-B residual GEMM (28.6 -> 31.6 us with 7 stages):
  // there the fp32 residual read by the epilogue is twice the A traffic and sits on the critical path
  int st = (EPI == HMA_EPI_RESID_F32 && stat) ? 4 : kMaxStages;
  while (st > 2 && smem_need<BN, EPI>(KB, stat, st) > (size_t)kSmemLimit) --st;
  return st;
}

template <int BN, int EPI, bool STAT, bool LN = false>
static int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmNtParams& p_in, cudaStream_t stream) {
  GemmNtParams p = p_in;
  p.stages = pick_stages<BN, EPI>(p.K / kBK, STAT);
  if (LN) {  // the LN variant keeps the tile's fp32 x on chip (kLnStash) and has ~7 KB of static shared memory
    while (p.stages > 2 && smem_need<BN, EPI>(p.K / kBK, STAT, p.stages) + kLnStash + kLnStatic > (size_t)kSmemLimit) --p.stages;
  }
  const size_t smem = smem_need<BN, EPI>(p.K / kBK, STAT, p.stages) + (LN ? kLnStash : 0);
  auto kern = gemm_nt_kernel<BN, EPI, STAT, LN>;
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();  // idempotent; racing threads set the same value
  constexpr int kDynLimit = kSmemLimit - (LN ? kLnStatic : 0);  // static + dynamic shared memory share the 227 KB
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynLimit));
    attr_done = true;
  }
  HMA_REQUIRE(smem <= (size_t)kDynLimit, "gemm_nt: shared memory request %zu too large", smem);
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  int per_n = hma_host::sm_count() / n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > m_tiles) per_n = m_tiles;
  if (LN) {  // the two CTAs of a row tile (n_blk 0 and 1 = consecutive blocks) form a cluster
    // persistent grid = the clusters that can be resident together (a GPC with an odd number of free SMs strands one)
    static hma_host::PerDeviceFlag occ_flag;
    static int occ_clusters[64];
    bool& occ_done = occ_flag.get();
    const int dev = hma_host::current_device();
    if (!occ_done) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(hma_host::sm_count() / 2 * 2);
      cfg.blockDim = dim3(128 + 32 * EpiWarps<EPI>::value);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      HMA_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      occ_clusters[dev & 63] = n > 0 ? n : 1;
      if (getenv("HMA_B200_VERBOSE")) fprintf(stderr, "[hma] gemm_nt_ln: %d two-CTA clusters resident on %d SMs\n", n, hma_host::sm_count());
      occ_done = true;
    }
    if (per_n > occ_clusters[dev & 63]) per_n = occ_clusters[dev & 63];
  }
  const int grid = per_n * n_tiles;
  if (LN) {
    HMA_CHECK_CUDA(hma_host::launch_pdl_cluster(kern, dim3(grid), dim3(128 + 32 * EpiWarps<EPI>::value), smem, 2, stream, tmA, tmB, p));
    return 0;
  }
  HMA_CHECK_CUDA(hma_host::launch_pdl(kern, dim3(grid), dim3(128 + 32 * EpiWarps<EPI>::value), smem, stream, tmA, tmB, p));
  return 0;
}

template <int EPI>
static int dispatch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmNtParams& p, int bn,
                       cudaStream_t stream) {
  const int KB = p.K / kBK;
  if constexpr (EpiWarps<EPI>::value == 8) {  // 64-wide tiles: one 32-column chunk per epilogue warp
    if (bn == 64) {
      const bool stat = smem_need<64, EPI>(KB, true, 2) <= (size_t)kSmemLimit;
      return stat ? launch_nt<64, EPI, true>(tmA, tmB, p, stream) : launch_nt<64, EPI, false>(tmA, tmB, p, stream);
    }
  }
  if (bn == 256) {
    const bool stat = smem_need<256, EPI>(KB, true, 2) <= (size_t)kSmemLimit;
    return stat ? launch_nt<256, EPI, true>(tmA, tmB, p, stream) : launch_nt<256, EPI, false>(tmA, tmB, p, stream);
  }
  const bool stat = smem_need<128, EPI>(KB, true, 2) <= (size_t)kSmemLimit;
  return stat ? launch_nt<128, EPI, true>(tmA, tmB, p, stream) : launch_nt<128, EPI, false>(tmA, tmB, p, stream);
}

}  // namespace hma

static int gemm_nt_impl(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                        int epi, void* out, long long ldo, void* out2, long long ldo2, const float* bias,
                        const float* resid, long long ldr, const void* aux, long long ldaux, float alpha,
                        float* colsum, float* rowdot, int conv_cin, int conv_wp, void* stream_,
                        const hma::GemmNtParams* ln = nullptr) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M == 0) return 0;
  HMA_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_nt: bad shape M=%d N=%d K=%d", M, N, K);
  HMA_REQUIRE(K % kBK == 0, "gemm_nt: K=%d must be a multiple of 64", K);
  HMA_REQUIRE(N % 128 == 0, "gemm_nt: N=%d must be a multiple of 128", N);
  HMA_REQUIRE(out != nullptr, "gemm_nt: out is null");
  HMA_REQUIRE(ldo % 8 == 0 && ldo2 % 8 == 0 && ldr % 4 == 0 && ldaux % 4 == 0, "gemm_nt: leading dimensions must keep rows 16-byte aligned");
  if (ln != nullptr) {  // (argument checks come before anything that needs the driver)
    HMA_REQUIRE(epi == HMA_EPI_RESID_F32 && N == 256 && out2 == nullptr && colsum == nullptr,
                "gemm_nt_ln: the LayerNorm epilogue needs the fp32 residual epilogue and N == 256 (a row = one 2-CTA cluster)");
    HMA_REQUIRE(ln->ln_mode == 1 || ln->ln_mode == 2, "gemm_nt_ln: bad LayerNorm mode %d", ln->ln_mode);
    HMA_REQUIRE(ln->ln_mode != 1 || (ln->ln_gamma && ln->ln_beta), "gemm_nt_ln: affine mode needs gamma/beta");
    HMA_REQUIRE(ln->ln_mode != 2 || (ln->ln_mod && ln->ln_rpg > 0), "gemm_nt_ln: modulate mode needs shift/scale");
    HMA_REQUIRE(ln->ln_out != nullptr && ln->ln_ldo % 4 == 0, "gemm_nt_ln: bad LayerNorm output");
  }
  int bn = (N % 256 == 0 && N >= 512) ? 256 : 128;
  // Few rows (decode / sampler shapes, M <= a few hundred): 256-wide tiles would occupy a fraction of the SMs and each
  // CTA's k-loop is latency-bound, so halve the tile width to double the CTAs in flight.
  if (bn == 256 && (long long)((M + kBM - 1) / kBM) * (N / 256) * 2 <= hma_host::sm_count()) bn = 128;
  // dGELU: the epilogue (3 loads, 2 MUFU ops, ~16 instructions per element) is latency-bound with two warps per scheduler;
  // sixteen epilogue warps need their fp32 staging buffers (74 KB), which only fits beside a stationary B with 128-wide tiles
  // Very few rows (the diffusion sampler's <= 512-row GEMMs, the one-frame decode passes at small batch): even 128-wide tiles
  // leave most SMs idle and every CTA's k-loop (K = 1024: 16 k-blocks) is a latency chain, so go to 64-wide tiles — 4x the
  // CTAs of the 256-wide choice, each streaming 24 KB instead of 48 KB per k-block and issuing N = 64 MMAs (49 vs 127 cycles).
  if (bn == 128 && conv_cin == 0 && (epi == HMA_EPI_BF16 || epi == HMA_EPI_RESID_F32 || epi == HMA_EPI_SILU_BF16) &&
      (long long)((M + kBM - 1) / kBM) * (N / 128) * 2 <= hma_host::sm_count())
    bn = 64;
  if (ln != nullptr) bn = 128;
  if (epi == HMA_EPI_DGELU_BF16) bn = 128;  // (the same change for the forward GELU epilogue measured worse: 52.7 vs 50.3 us)
  CUtensorMap tmA, tmB;
  int rc = hma_host::make_tmap_bf16_2d(&tmA, A, (uint64_t)(conv_cin > 0 ? conv_cin : K), (uint64_t)M, (uint64_t)lda * 2, kBK, kBM);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBK, bn);
  if (rc) return rc;
  GemmNtParams p;
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldo = ldo; p.out2 = out2; p.ldo2 = ldo2;
  p.bias = bias; p.resid = resid; p.ldr = ldr;
  p.aux = static_cast<const __nv_bfloat16*>(aux); p.ldaux = ldaux;
  p.alpha = alpha;
  p.colsum = colsum;
  p.rowdot = rowdot;
  p.conv_kb_per_tap = conv_cin > 0 ? conv_cin / kBK : 0;
  p.conv_wp = conv_wp;
  p.ln_mode = 0; p.ln_gamma = p.ln_beta = p.ln_mod = nullptr; p.ln_rpg = 0; p.ln_eps = 0.f; p.ln_out = nullptr; p.ln_ldo = 0;
  p.ln_stats = nullptr;
  HMA_REQUIRE(rowdot == nullptr || (epi == HMA_EPI_BF16 && aux != nullptr && ldaux % 8 == 0),
              "gemm_nt: rowdot needs the plain bf16 epilogue and a bf16 aux matrix with 16-byte aligned rows");
  HMA_REQUIRE(colsum == nullptr || epi == HMA_EPI_DGELU_BF16 || epi == HMA_EPI_DSILU_BF16 ||
                  (epi == HMA_EPI_RESID_F32 && out2 != nullptr),
              "gemm_nt: colsum is produced by the d-activation epilogues and by the residual epilogue with a bf16 copy");
  if (ln != nullptr) {
    p.ln_mode = ln->ln_mode; p.ln_gamma = ln->ln_gamma; p.ln_beta = ln->ln_beta; p.ln_mod = ln->ln_mod; p.ln_rpg = ln->ln_rpg;
    p.ln_eps = ln->ln_eps; p.ln_out = ln->ln_out; p.ln_ldo = ln->ln_ldo; p.ln_stats = ln->ln_stats;
    const bool stat = smem_need<128, HMA_EPI_RESID_F32>(p.K / kBK, true, 3) + kLnStash + kLnStatic <= (size_t)kSmemLimit;
    return stat ? launch_nt<128, HMA_EPI_RESID_F32, true, true>(tmA, tmB, p, stream)
                : launch_nt<128, HMA_EPI_RESID_F32, false, true>(tmA, tmB, p, stream);
  }
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

extern "C" int hma_gemm_nt(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                           int epi, void* out, long long ldo, void* out2, long long ldo2, const float* bias,
                           const float* resid, long long ldr, const void* aux, long long ldaux, float alpha,
                           float* colsum, float* rowdot, void* stream_) {
  return gemm_nt_impl(A, lda, B, ldb, M, N, K, epi, out, ldo, out2, ldo2, bias, resid, ldr, aux, ldaux, alpha, colsum, rowdot, 0, 0,
                      stream_);
}

// out = resid + alpha * A B^T + bias (fp32, N == 256) AND ln_out = bf16(LayerNorm(out)) from the same epilogue: the pre-norm of
// the NEXT stage (st_transformer.py:85-86,112 norm1 / norm2: mode 1, eps 1e-5, affine; st_mask_git.py:66-76 ModulateLayer's
// LayerNorm(elementwise_affine=False, eps 1e-6) * (1 + scale) + shift per frame: mode 2) without a separate pass over the
// residual stream. stats (optional) = (mean, rstd) per row, what hma_ln_bwd reads.
extern "C" int hma_gemm_nt_ln(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, void* out,
                              long long ldo, const float* bias, const float* resid, long long ldr, float alpha, int ln_mode,
                              const float* gamma, const float* beta, const float* mod, int rows_per_group, float eps,
                              void* ln_out, long long ld_ln, float* stats, void* stream_) {
  hma::GemmNtParams ln;
  ln.ln_mode = ln_mode; ln.ln_gamma = gamma; ln.ln_beta = beta; ln.ln_mod = mod; ln.ln_rpg = rows_per_group; ln.ln_eps = eps;
  ln.ln_out = static_cast<__nv_bfloat16*>(ln_out); ln.ln_ldo = ld_ln; ln.ln_stats = stats;
  return gemm_nt_impl(A, lda, B, ldb, M, N, K, HMA_EPI_RESID_F32, out, ldo, nullptr, 0, bias, resid, ldr, nullptr, 0, alpha,
                      nullptr, nullptr, 0, 0, stream_, &ln);
}

// 3x3 convolution, stride 1, zero padding 1 (nn.Conv2d(k=3, padding=1): external/magvit2 improved_model.py:27-28,135,160,228)
// over zero-bordered NHWC images: X is bf16 [images * (H+2) * (W+2), Cin] with every border pixel zero, Wt is bf16
// [Cout, 9 * Cin] (k index = (ky * 3 + kx) * Cin + ci), out fp32 [images * (H+2) * (W+2), Cout] = resid + conv + bias on the
// interior pixels (border rows of `out` receive values that mean nothing: the next stage re-zeroes them).
extern "C" int hma_conv3x3_nhwc(const void* X, long long ldx, const void* Wt, long long ldw, int rows, int Cin, int Cout,
                                int padded_width, void* out, long long ldo, const float* bias, const float* resid,
                                long long ldr, void* stream_) {
  HMA_REQUIRE(Cin > 0 && Cin % hma::kBK == 0, "conv3x3: Cin=%d must be a multiple of 64", Cin);
  HMA_REQUIRE(padded_width >= 3, "conv3x3: bad padded width %d", padded_width);
  return gemm_nt_impl(X, ldx, Wt, ldw, rows, Cout, 9 * Cin, HMA_EPI_RESID_F32, out, ldo, nullptr, 0, bias, resid, ldr, nullptr, 0,
                      1.0f, nullptr, nullptr, Cin, padded_width, stream_);
}
