// Per-head LayerNorm of q and k ("qk_norm", attention.py:32-35,47-52 / 143-148): after the QKV projection every
// 32-channel head slice of q and of k is normalised with ONE shared affine LayerNorm(head_dim, eps 1e-5), in fp32,
// and cast back to bf16; v passes through. Forward writes [LN(q) | LN(k) | v] to a second matrix (the attention
// kernels address q, k, v as column ranges of one matrix); backward turns the gradient w.r.t. that matrix into the
// gradient w.r.t. the raw projection output in place and accumulates d(gamma), d(beta).
// HBM-bound row-wise kernels: one warp per token row, lane l owns the 16 contiguous channels [16 l, 16 l + 16) of the
// 512 q|k channels, i.e. half a head slice (the two lanes of a slice combine with one shuffle), plus 8 channels of v.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kQC = 256;  // channels of q (= of k, = of v): 8 heads x 32

__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[2 * i] = bf16_lo(w[i]); f[2 * i + 1] = bf16_hi(w[i]); }
}
__device__ __forceinline__ void pack16(const float (&f)[16], uint4& a, uint4& b) {
  a = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  b = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
}

// mean / rstd of the 32-channel slice shared by lanes (l, l ^ 1)
__device__ __forceinline__ void slice_stats(const float (&x)[16], float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  mean = s * (1.0f / 32.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  rstd = rsqrtf(q * (1.0f / 32.0f) + eps);
}

__global__ void __launch_bounds__(256) qk_norm_fwd_kernel(const __nv_bfloat16* qkv, long long ld, int rows, const float* gamma,
                                                          const float* beta, float eps, __nv_bfloat16* out, long long ldo) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const __nv_bfloat16* src = qkv + row * ld;
  __nv_bfloat16* dst = out + row * ldo;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + lane * 16);
  float x[16];
  unpack16(s4[0], s4[1], x);
  *reinterpret_cast<uint4*>(dst + 2 * kQC + lane * 8) = *reinterpret_cast<const uint4*>(src + 2 * kQC + lane * 8);  // v
  float mean, rstd;
  slice_stats(x, eps, mean, rstd);
  const int c0 = (lane & 1) * 16;
  float y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] = fmaf((x[i] - mean) * rstd, __ldg(gamma + c0 + i), __ldg(beta + c0 + i));
  uint4 a, b;
  pack16(y, a, b);
  uint4* d4 = reinterpret_cast<uint4*>(dst + lane * 16);
  d4[0] = a;
  d4[1] = b;
}

// dqkv: gradient w.r.t. [LN(q) | LN(k) | v] on entry, w.r.t. the raw q | k | v on exit (v columns untouched).
__global__ void __launch_bounds__(256) qk_norm_bwd_kernel(const __nv_bfloat16* qkv, long long ld, int rows, int rows_per_cta,
                                                          const float* gamma, float eps, __nv_bfloat16* dqkv, long long ldd,
                                                          float* dgamma, float* dbeta) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (lane & 1) * 16;
  float g[16], acc_g[16], acc_b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { g[i] = __ldg(gamma + c0 + i); acc_g[i] = 0.f; acc_b[i] = 0.f; }
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  for (int i = warp; i < rows_per_cta; i += 8) {
    const long long row = r0 + i;
    if (row >= rows) break;
    const uint4* s4 = reinterpret_cast<const uint4*>(qkv + row * ld + lane * 16);
    uint4* d4 = reinterpret_cast<uint4*>(dqkv + row * ldd + lane * 16);
    float x[16], dy[16];
    unpack16(s4[0], s4[1], x);
    unpack16(d4[0], d4[1], dy);
    float mean, rstd;
    slice_stats(x, eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f, xh[16], gg[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      xh[k] = (x[k] - mean) * rstd;
      gg[k] = dy[k] * g[k];
      s1 += gg[k];
      s2 = fmaf(gg[k], xh[k], s2);
      acc_g[k] = fmaf(dy[k], xh[k], acc_g[k]);
      acc_b[k] += dy[k];
    }
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    s1 *= (1.0f / 32.0f);
    s2 *= (1.0f / 32.0f);
    float dx[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) dx[k] = rstd * (gg[k] - s1 - xh[k] * s2);
    uint4 a, b;
    pack16(dx, a, b);
    d4[0] = a;
    d4[1] = b;
  }
  // lanes of equal parity own the same 16 slice positions: reduce over them, then over the CTA's warps
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) {
      acc_g[k] += __shfl_xor_sync(0xffffffffu, acc_g[k], o);
      acc_b[k] += __shfl_xor_sync(0xffffffffu, acc_b[k], o);
    }
  }
  if (lane < 2) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      red[warp][lane * 16 + k] = acc_g[k];
      red[warp][32 + lane * 16 + k] = acc_b[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (threadIdx.x < 32) atomicAdd(dgamma + threadIdx.x, s);
    else atomicAdd(dbeta + threadIdx.x - 32, s);
  }
}

}  // namespace hma

extern "C" int hma_qk_norm_fwd(const void* qkv, long long ld, int rows, const float* gamma, const float* beta, float eps,
                               void* out, long long ldo, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ld >= 3 * kQC && ldo >= 3 * kQC, "qk_norm_fwd: needs [rows, >= 768] matrices, 16-byte aligned rows");
  HMA_REQUIRE(gamma != nullptr && beta != nullptr, "qk_norm_fwd: null affine parameters");
  HMA_CHECK_CUDA(hma_host::launch_pdl(qk_norm_fwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream_),
                                      static_cast<const __nv_bfloat16*>(qkv), ld, rows, gamma, beta, eps,
                                      static_cast<__nv_bfloat16*>(out), ldo));
  return 0;
}

extern "C" int hma_qk_norm_bwd(const void* qkv, long long ld, int rows, const float* gamma, float eps, void* dqkv,
                               long long ldd, float* dgamma, float* dbeta, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(ld % 8 == 0 && ldd % 8 == 0 && ld >= 3 * kQC && ldd >= 3 * kQC, "qk_norm_bwd: needs [rows, >= 768] matrices, 16-byte aligned rows");
  HMA_REQUIRE(gamma != nullptr && dgamma != nullptr && dbeta != nullptr, "qk_norm_bwd: null parameter pointers");
  const int rpc = 64;
  HMA_CHECK_CUDA(hma_host::launch_pdl(qk_norm_bwd_kernel, dim3((rows + rpc - 1) / rpc), dim3(256), 0, static_cast<cudaStream_t>(stream_),
                                      static_cast<const __nv_bfloat16*>(qkv), ld, rows, rpc, gamma, eps,
                                      static_cast<__nv_bfloat16*>(dqkv), ldd, dgamma, dbeta));
  return 0;
}
