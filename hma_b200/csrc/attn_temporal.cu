// Causal attention over the T frames of each spatial slot (reference: attention.py:37-61 with
// causal=True, called from st_transformer.py:111; no pre-norm). T is 4..64, head_dim 32: each
// sequence is a handful of 32-wide dot products, 0.3 % of the model FLOPs, so this stage is
// bound by the bytes of qkv / out, not by math, and runs on the CUDA cores with the whole
// working set of a CTA staged in shared memory. The (B,T,n,C) residual layout is read in place:
// the T rows of one sequence are n*3C elements apart, so the reference's "(B T) S C -> (B S) T C"
// transposes (st_transformer.py:89,113) never materialise.
//
// One CTA = one sample b, SC consecutive slots s, all heads, all T frames.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct TemporalParams {
  int B, T, n, heads;
  int SC;  // slots per CTA
  float scale;
  const __nv_bfloat16* qkv;  // [B*T*n, 3C]
  long long ld_qkv;
  int q_col, k_col, v_col;
  __nv_bfloat16* out;  // fwd: [B*T*n, C]
  long long ldo;
  const __nv_bfloat16* dout;  // bwd: [B*T*n, C]
  long long ld_dout;
  __nv_bfloat16* dqkv;  // bwd: [B*T*n, 3C]
  long long ld_dqkv;
};

__device__ __forceinline__ void load_row32(const __nv_bfloat16* src, float (&dst)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = s4[q];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dst[q * 8 + 2 * j] = bf16_lo(w[j]);
      dst[q * 8 + 2 * j + 1] = bf16_hi(w[j]);
    }
  }
}
__device__ __forceinline__ float dot_row32(const __nv_bfloat16* src, const float (&a)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = s4[q];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fmaf(bf16_lo(w[j]), a[q * 8 + 2 * j], acc);
      acc = fmaf(bf16_hi(w[j]), a[q * 8 + 2 * j + 1], acc);
    }
  }
  return acc;
}
__device__ __forceinline__ void axpy_row32(const __nv_bfloat16* src, float alpha, float (&acc)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = s4[q];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[q * 8 + 2 * j] = fmaf(alpha, bf16_lo(w[j]), acc[q * 8 + 2 * j]);
      acc[q * 8 + 2 * j + 1] = fmaf(alpha, bf16_hi(w[j]), acc[q * 8 + 2 * j + 1]);
    }
  }
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d4[q] = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
}

// Stage [T][SC][width] bf16 rows (width = 3C or C) of sample b, slots s0.. into shared memory.
__device__ __forceinline__ void stage_rows(const __nv_bfloat16* g, long long ld, int width, int b, int T, int n, int s0,
                                           int sc, int SC, __nv_bfloat16* smem) {
  const int vec_per_row = width / 8;
  const int total = T * sc * vec_per_row;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int v = i % vec_per_row;
    const int rs = i / vec_per_row;
    const int sl = rs % sc;
    const int t = rs / sc;
    const size_t row = (size_t)(b * T + t) * n + s0 + sl;
    const uint4 val = *reinterpret_cast<const uint4*>(g + row * ld + v * 8);
    *reinterpret_cast<uint4*>(smem + ((size_t)(t * SC + sl) * width) + v * 8) = val;
  }
}

__global__ void __launch_bounds__(256) attn_temporal_fwd_kernel(const TemporalParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sqkv = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  const int C = p.heads * 32, W = 3 * C;
  const int chunks = (p.n + p.SC - 1) / p.SC;
  const int b = blockIdx.x / chunks;
  const int s0 = (blockIdx.x % chunks) * p.SC;
  const int sc = min(p.SC, p.n - s0);
  // the q/k/v column offsets are relative to the staged row, which starts at qkv column 0
  stage_rows(p.qkv, p.ld_qkv, W, b, p.T, p.n, s0, sc, p.SC, sqkv);
  __syncthreads();
  const int items = sc * p.heads * p.T;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tq = it % p.T;
    const int h = (it / p.T) % p.heads;
    const int sl = it / (p.T * p.heads);
    float q[32], acc[32];
    load_row32(sqkv + (size_t)(tq * p.SC + sl) * W + p.q_col + h * 32, q);
#pragma unroll
    for (int j = 0; j < 32; ++j) { q[j] *= p.scale; acc[j] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int tk = 0; tk <= tq; ++tk) {
      const __nv_bfloat16* krow = sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32;
      const float s = dot_row32(krow, q);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn);
      const float pe = __expf(s - mn);
      l = l * corr + pe;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] *= corr;
      axpy_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.v_col + h * 32, pe, acc);
      m = mn;
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= inv;
    const size_t row = (size_t)(b * p.T + tq) * p.n + s0 + sl;
    store_row32(p.out + row * p.ldo + h * 32, acc);
  }
}

// Backward: recomputes the T x T probabilities. Phase 1 (one thread per query row) produces dq and
// the row statistics (max, sum, delta = dO.O); phase 2 (one thread per key row) produces dk, dv.
__global__ void __launch_bounds__(256) attn_temporal_bwd_kernel(const TemporalParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int C = p.heads * 32, W = 3 * C;
  __nv_bfloat16* sqkv = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sdo = sqkv + (size_t)p.T * p.SC * W;
  float* sstat = reinterpret_cast<float*>(sdo + (size_t)p.T * p.SC * C);  // [items][3]: m, 1/l, delta
  const int chunks = (p.n + p.SC - 1) / p.SC;
  const int b = blockIdx.x / chunks;
  const int s0 = (blockIdx.x % chunks) * p.SC;
  const int sc = min(p.SC, p.n - s0);
  stage_rows(p.qkv, p.ld_qkv, W, b, p.T, p.n, s0, sc, p.SC, sqkv);
  stage_rows(p.dout, p.ld_dout, C, b, p.T, p.n, s0, sc, p.SC, sdo);
  __syncthreads();
  const int items = sc * p.heads * p.T;
  // ---------------- phase 1: per query row
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tq = it % p.T;
    const int h = (it / p.T) % p.heads;
    const int sl = it / (p.T * p.heads);
    float q[32], dO[32];
    load_row32(sqkv + (size_t)(tq * p.SC + sl) * W + p.q_col + h * 32, q);
    load_row32(sdo + (size_t)(tq * p.SC + sl) * C + h * 32, dO);
#pragma unroll
    for (int j = 0; j < 32; ++j) q[j] *= p.scale;
    float m = -INFINITY;
    for (int tk = 0; tk <= tq; ++tk)
      m = fmaxf(m, dot_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32, q));
    float l = 0.f, delta = 0.f;
    for (int tk = 0; tk <= tq; ++tk) {
      const float s = dot_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32, q);
      const float pe = __expf(s - m);
      l += pe;
      delta += pe * dot_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.v_col + h * 32, dO);
    }
    const float inv = 1.0f / l;
    delta *= inv;
    float dq[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) dq[j] = 0.f;
    for (int tk = 0; tk <= tq; ++tk) {
      const float s = dot_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32, q);
      const float pr = __expf(s - m) * inv;
      const float dP = dot_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.v_col + h * 32, dO);
      const float dS = pr * (dP - delta) * p.scale;
      axpy_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32, dS, dq);
    }
    sstat[it * 3 + 0] = m;
    sstat[it * 3 + 1] = inv;
    sstat[it * 3 + 2] = delta;
    const size_t row = (size_t)(b * p.T + tq) * p.n + s0 + sl;
    store_row32(p.dqkv + row * p.ld_dqkv + p.q_col + h * 32, dq);
  }
  __syncthreads();
  // ---------------- phase 2: per key row
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tk = it % p.T;
    const int h = (it / p.T) % p.heads;
    const int sl = it / (p.T * p.heads);
    float k[32], v[32], dk[32], dv[32];
    load_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.k_col + h * 32, k);
    load_row32(sqkv + (size_t)(tk * p.SC + sl) * W + p.v_col + h * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) { dk[j] = 0.f; dv[j] = 0.f; }
    for (int tq = tk; tq < p.T; ++tq) {
      const int iq = (sl * p.heads + h) * p.T + tq;
      const __nv_bfloat16* qrow = sqkv + (size_t)(tq * p.SC + sl) * W + p.q_col + h * 32;
      const __nv_bfloat16* dorow = sdo + (size_t)(tq * p.SC + sl) * C + h * 32;
      const float s = dot_row32(qrow, k) * p.scale;
      const float pr = __expf(s - sstat[iq * 3 + 0]) * sstat[iq * 3 + 1];
      const float dP = dot_row32(dorow, v);
      const float dS = pr * (dP - sstat[iq * 3 + 2]) * p.scale;
      axpy_row32(qrow, dS, dk);
      axpy_row32(dorow, pr, dv);
    }
    const size_t row = (size_t)(b * p.T + tk) * p.n + s0 + sl;
    store_row32(p.dqkv + row * p.ld_dqkv + p.k_col + h * 32, dk);
    store_row32(p.dqkv + row * p.ld_dqkv + p.v_col + h * 32, dv);
  }
}

static int pick_sc(int T, int heads, bool bwd) {
  const int C = heads * 32;
  for (int sc = 4; sc >= 1; sc >>= 1) {
    size_t bytes = (size_t)T * sc * 3 * C * 2;
    if (bwd) bytes += (size_t)T * sc * C * 2 + (size_t)sc * heads * T * 12;
    if (bytes <= 200 * 1024) return sc;
  }
  return 0;
}

}  // namespace hma

extern "C" int hma_attn_temporal_fwd(const void* qkv, long long ld_qkv, int B, int T, int n, int heads, int q_col,
                                     int k_col, int v_col, float scale, void* out, long long ldo, void* stream_) {
  using namespace hma;
  if (B == 0) return 0;
  const int sc = pick_sc(T, heads, false);
  HMA_REQUIRE(sc > 0 && T >= 1, "attn_temporal: T=%d too long for the shared-memory staging", T);
  HMA_REQUIRE(q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 && ld_qkv % 8 == 0 && ldo % 8 == 0,
              "attn_temporal: 16-byte alignment required");
  HMA_REQUIRE(q_col + heads * 32 <= 3 * heads * 32 && k_col + heads * 32 <= 3 * heads * 32 &&
                  v_col + heads * 32 <= 3 * heads * 32, "attn_temporal: q/k/v must live inside one 3C-wide row");
  TemporalParams p{};
  p.B = B; p.T = T; p.n = n; p.heads = heads; p.SC = sc; p.scale = scale;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv); p.ld_qkv = ld_qkv;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  const size_t smem = (size_t)T * sc * 3 * heads * 32 * 2;
  static size_t attr = 0;
  if (smem > attr) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int chunks = (n + sc - 1) / sc;
  attn_temporal_fwd_kernel<<<B * chunks, 256, smem, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_attn_temporal_bwd(const void* qkv, long long ld_qkv, const void* dout, long long ld_dout, int B,
                                     int T, int n, int heads, int q_col, int k_col, int v_col, float scale,
                                     void* dqkv, long long ld_dqkv, void* stream_) {
  using namespace hma;
  if (B == 0) return 0;
  const int sc = pick_sc(T, heads, true);
  HMA_REQUIRE(sc > 0 && T >= 1, "attn_temporal_bwd: T=%d too long for the shared-memory staging", T);
  HMA_REQUIRE(ld_qkv % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "attn_temporal_bwd: 16-byte alignment required");
  TemporalParams p{};
  p.B = B; p.T = T; p.n = n; p.heads = heads; p.SC = sc; p.scale = scale;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv); p.ld_qkv = ld_qkv;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.ld_dout = ld_dout;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  const int C = heads * 32;
  const size_t smem = (size_t)T * sc * 3 * C * 2 + (size_t)T * sc * C * 2 + (size_t)sc * heads * T * 12;
  static size_t attr = 0;
  if (smem > attr) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(attn_temporal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int chunks = (n + sc - 1) / sc;
  attn_temporal_bwd_kernel<<<B * chunks, 256, smem, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
