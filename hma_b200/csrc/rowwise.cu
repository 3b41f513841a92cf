// Row-wise, HBM-bound kernels around the tensor-core contractions (C = d_model = 256 only):
//   * LayerNorm forward in three flavours, fp32 residual stream in -> bf16 GEMM operand out:
//       mode 0  plain cast                        (temporal attention input, st_transformer.py:111)
//       mode 1  affine LayerNorm, eps 1e-5        (norm1 / norm2, st_transformer.py:50,75,86,112)
//       mode 2  LayerNorm without affine, eps 1e-6, then x*(1+scale)+shift with per-(b,t) shift /
//               scale                              (ModulateLayer, st_mask_git.py:18-19,58,70-75)
//   * the matching backward, accumulating into the fp32 gradient stream and the parameter grads
//   * bf16 column sums (bias gradients), fp32 -> bf16 weight cast (+ transposed copy)
// One warp per row, 8 floats per lane, vector loads; every kernel streams each byte once.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kC = 256;

struct RowLoad {
  float v[8];
};
__device__ __forceinline__ RowLoad load_row_f32(const float* row, int lane) {
  RowLoad r;
  const float4 a = *reinterpret_cast<const float4*>(row + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ RowLoad load_row_bf16(const __nv_bfloat16* row, int lane) {
  RowLoad r;
  const uint2 a = *reinterpret_cast<const uint2*>(row + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(row + 128 + lane * 4);
  r.v[0] = bf16_lo(a.x); r.v[1] = bf16_hi(a.x); r.v[2] = bf16_lo(a.y); r.v[3] = bf16_hi(a.y);
  r.v[4] = bf16_lo(b.x); r.v[5] = bf16_hi(b.x); r.v[6] = bf16_lo(b.y); r.v[7] = bf16_hi(b.y);
  return r;
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* row, int lane, const float (&v)[8]) {
  *reinterpret_cast<uint2*>(row + lane * 4) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
  *reinterpret_cast<uint2*>(row + 128 + lane * 4) = make_uint2(pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void store_row_f32(float* row, int lane, const float (&v)[8]) {
  *reinterpret_cast<float4*>(row + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(row + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// column owned by (lane, j)
__device__ __forceinline__ int col_of(int lane, int j) { return (j < 4 ? 0 : 128) + lane * 4 + (j & 3); }

struct LnParams {
  const float* x;
  long long ldx;
  int rows;
  int mode;
  const float* gamma;
  const float* beta;
  const float* mod;  // [groups, 2C]: shift | scale
  int rows_per_group;
  float eps;
  __nv_bfloat16* y;
  long long ldy;
  float* stats;  // [rows, 2] mean, rstd (may be null)
  int src_group, dst_group;  // output row r reads input row (r / dst_group) * src_group + r % dst_group
};

// Rows are processed four at a time per warp (all eight 16-byte loads of a lane are issued before any reduction),
// by a grid of at most a few CTAs per SM striding over the row quads: one row per warp with a fresh CTA per
// 8 rows reached only ~60 % of the copy bandwidth.
constexpr int kLnRowsPerWarp = 4;

__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gam[8], bet[8];
  if (p.mode == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { gam[j] = __ldg(p.gamma + col_of(lane, j)); bet[j] = __ldg(p.beta + col_of(lane, j)); }
  }
  const int quads = (p.rows + kLnRowsPerWarp - 1) / kLnRowsPerWarp;
  for (int q = blockIdx.x * 8 + warp; q < quads; q += gridDim.x * 8) {
    const int row0 = q * kLnRowsPerWarp;
    RowLoad r[kLnRowsPerWarp];
#pragma unroll
    for (int u = 0; u < kLnRowsPerWarp; ++u) {
      const int row = min(row0 + u, p.rows - 1);
      const size_t src_row = p.dst_group > 0 ? (size_t)(row / p.dst_group) * p.src_group + (row % p.dst_group) : (size_t)row;
      r[u] = load_row_f32(p.x + src_row * p.ldx, lane);
    }
#pragma unroll
    for (int u = 0; u < kLnRowsPerWarp; ++u) {
      const int row = row0 + u;
      if (row >= p.rows) break;
      float out[8];
      if (p.mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = r[u].v[j];
      } else {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += r[u].v[j];
        const float mean = warp_sum(s) * (1.0f / kC);
        float qq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = r[u].v[j] - mean; qq = fmaf(d, d, qq); }
        const float rstd = rsqrtf(warp_sum(qq) * (1.0f / kC) + p.eps);
        if (p.stats != nullptr && lane == 0) {
          p.stats[(size_t)row * 2] = mean;
          p.stats[(size_t)row * 2 + 1] = rstd;
        }
        if (p.mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) out[j] = fmaf((r[u].v[j] - mean) * rstd, gam[j], bet[j]);
        } else {
          const float* m = p.mod + (size_t)(row / p.rows_per_group) * 2 * kC;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = col_of(lane, j);
            out[j] = (r[u].v[j] - mean) * rstd * (1.0f + __ldg(m + kC + c)) + __ldg(m + c);
          }
        }
      }
      store_row_bf16(p.y + (size_t)row * p.ldy, lane, out);
    }
  }
}

struct LnBwdParams {
  const __nv_bfloat16* dy;
  long long lddy;
  const float* x;
  long long ldx;
  const float* stats;
  int rows;
  int mode;  // 1 or 2
  const float* gamma;
  const float* mod;
  int rows_per_group;
  int rows_per_cta;
  float* dx;  // accumulated in place
  long long lddx;
  float* dgamma;  // mode 1: [C] dgamma, [C] dbeta (separate pointers)
  float* dbeta;
  float* dmod;  // mode 2: [groups, 2C] dshift | dscale
  __nv_bfloat16* dy_next;  // optional: bf16 copy of the updated dx (the next stage's GEMM operand)
  float* colsum_next;      // optional: += column sums of dy_next (the next stage's bias gradient)
};

__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][2 * kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * p.rows_per_cta;
  float acc_w[8], acc_b[8];  // d(gamma|scale), d(beta|shift) partials for this lane's 8 columns
  float acc_c[8];            // column sums of the bf16 copy of the new dx
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc_w[j] = 0.f; acc_b[j] = 0.f; acc_c[j] = 0.f; }
  float mult[8];
  if (p.mode == 0) {  // Identity "norm" (qk_norm=True, st_transformer.py:50,75): dx += dy, plus the bf16 copy / column sums
#pragma unroll
    for (int j = 0; j < 8; ++j) mult[j] = 0.f;
  } else if (p.mode == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) mult[j] = __ldg(p.gamma + col_of(lane, j));
  } else {
    const float* m = p.mod + (size_t)(r0 / p.rows_per_group) * 2 * kC;
#pragma unroll
    for (int j = 0; j < 8; ++j) mult[j] = 1.0f + __ldg(m + kC + col_of(lane, j));
  }
  for (int i = warp; i < p.rows_per_cta; i += 8) {
    const int row = r0 + i;
    if (row >= p.rows) break;
    const RowLoad dy = load_row_bf16(p.dy + (size_t)row * p.lddy, lane);
    RowLoad dx = load_row_f32(p.dx + (size_t)row * p.lddx, lane);
    float o[8];
    if (p.mode != 0) {
    const RowLoad x = load_row_f32(p.x + (size_t)row * p.ldx, lane);
    const float mean = __ldg(p.stats + (size_t)row * 2);
    const float rstd = __ldg(p.stats + (size_t)row * 2 + 1);
    float xh[8], g[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = (x.v[j] - mean) * rstd;
      g[j] = dy.v[j] * mult[j];
      s1 += g[j];
      s2 = fmaf(g[j], xh[j], s2);
      acc_w[j] = fmaf(dy.v[j], xh[j], acc_w[j]);
      acc_b[j] += dy.v[j];
    }
    s1 = warp_sum(s1) * (1.0f / kC);
    s2 = warp_sum(s2) * (1.0f / kC);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = dx.v[j] + rstd * (g[j] - s1 - xh[j] * s2);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = dx.v[j] + dy.v[j];
    }
    store_row_f32(p.dx + (size_t)row * p.lddx, lane, o);
    if (p.dy_next != nullptr) {
      const uint2 a = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
      const uint2 b = make_uint2(pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      *reinterpret_cast<uint2*>(p.dy_next + (size_t)row * kC + lane * 4) = a;
      *reinterpret_cast<uint2*>(p.dy_next + (size_t)row * kC + 128 + lane * 4) = b;
      acc_c[0] += bf16_lo(a.x); acc_c[1] += bf16_hi(a.x); acc_c[2] += bf16_lo(a.y); acc_c[3] += bf16_hi(a.y);
      acc_c[4] += bf16_lo(b.x); acc_c[5] += bf16_hi(b.x); acc_c[6] += bf16_lo(b.y); acc_c[7] += bf16_hi(b.y);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[warp][col_of(lane, j)] = acc_b[j];        // beta | shift first (matches the [shift|scale] layout)
    red[warp][kC + col_of(lane, j)] = acc_w[j];   // gamma | scale
  }
  __syncthreads();
  for (int c = threadIdx.x; p.mode != 0 && c < 2 * kC; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    if (p.mode == 1) {
      if (c < kC) atomicAdd(p.dbeta + c, s); else atomicAdd(p.dgamma + (c - kC), s);
    } else {
      atomicAdd(p.dmod + (size_t)(r0 / p.rows_per_group) * 2 * kC + c, s);
    }
  }
  if (p.colsum_next != nullptr) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][col_of(lane, j)] = acc_c[j];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(p.colsum_next + threadIdx.x, s);
  }
}

// out[c] += sum_r G[r, c]; one CTA per 128-row slab, 4 columns per thread (C <= 1024).
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* G, long long ld, int rows, int C,
                                                     float* out) {
  const int r0 = blockIdx.x * 128;
  const int r1 = min(rows, r0 + 128);
  const int c = threadIdx.x * 4;
  if (c >= C) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const uint2 v = *reinterpret_cast<const uint2*>(G + (size_t)r * ld + c);
    a0 += bf16_lo(v.x); a1 += bf16_hi(v.x); a2 += bf16_lo(v.y); a3 += bf16_hi(v.y);
  }
  atomicAdd(out + c, a0); atomicAdd(out + c + 1, a1); atomicAdd(out + c + 2, a2); atomicAdd(out + c + 3, a3);
}

// W fp32 [R, Cc] -> Wb bf16 [R, Cc] (optional) and Wt bf16 [Cc, R] (optional); 32x32 tiles.
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* W, int R, int Cc, __nv_bfloat16* Wb,
                                                             __nv_bfloat16* Wt, float alpha) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = W[(size_t)r * Cc + c] * alpha;
      if (Wb != nullptr) Wb[(size_t)r * Cc + c] = __float2bfloat16(v);
    }
    tile[i][tx] = v;
  }
  if (Wt == nullptr) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < Cc) Wt[(size_t)c * R + r] = __float2bfloat16(tile[tx][i]);
  }
}

}  // namespace hma

extern "C" int hma_ln_fwd(const float* x, long long ldx, int rows, int mode, const float* gamma, const float* beta,
                          const float* mod, int rows_per_group, float eps, void* y, long long ldy, float* stats,
                          int src_group, int dst_group, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(mode >= 0 && mode <= 2, "ln_fwd: bad mode %d", mode);
  HMA_REQUIRE(mode != 1 || (gamma && beta), "ln_fwd: affine mode needs gamma/beta");
  HMA_REQUIRE(mode != 2 || (mod && rows_per_group > 0), "ln_fwd: modulate mode needs shift/scale");
  HMA_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "ln_fwd: rows must be 16-byte aligned");
  HMA_REQUIRE((src_group == 0) == (dst_group == 0) && dst_group <= (src_group ? src_group : dst_group),
              "ln_fwd: bad row remap %d -> %d", src_group, dst_group);
  LnParams p{x, ldx, rows, mode, gamma, beta, mod, rows_per_group, eps, static_cast<__nv_bfloat16*>(y), ldy, stats,
             src_group, dst_group};
    int ln_grid = (rows + 8 * kLnRowsPerWarp - 1) / (8 * kLnRowsPerWarp);
  const int ln_cap = hma_host::sm_count() * 8;
  if (ln_grid > ln_cap) ln_grid = ln_cap;
  HMA_CHECK_CUDA(hma_host::launch_pdl(ln_fwd_kernel, dim3(ln_grid), dim3(256), 0, static_cast<cudaStream_t>(stream_), p));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_ln_bwd(const void* dy, long long lddy, const float* x, long long ldx, const float* stats, int rows,
                          int mode, const float* gamma, const float* mod, int rows_per_group, float* dx,
                          long long lddx, float* dgamma, float* dbeta, float* dmod, void* dy_next,
                          float* colsum_next, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(mode == 0 || mode == 1 || mode == 2, "ln_bwd: bad mode %d", mode);
  HMA_REQUIRE(mode != 1 || (gamma && dgamma && dbeta), "ln_bwd: affine mode needs gamma and grad buffers");
  HMA_REQUIRE(mode != 2 || (mod && dmod && rows_per_group > 0 && rows_per_group % 16 == 0),
              "ln_bwd: modulate mode needs shift/scale and rows_per_group %% 16 == 0");
  int rpc = 32;
  if (mode == 2 && rows_per_group % 32 != 0) rpc = 16;
  LnBwdParams p{static_cast<const __nv_bfloat16*>(dy), lddy, x, ldx, stats, rows, mode, gamma, mod,
                rows_per_group > 0 ? rows_per_group : rows, rpc, dx, lddx, dgamma, dbeta, dmod,
                static_cast<__nv_bfloat16*>(dy_next), colsum_next};
  HMA_REQUIRE(colsum_next == nullptr || dy_next != nullptr, "ln_bwd: colsum_next needs dy_next");
  HMA_CHECK_CUDA(hma_host::launch_pdl(ln_bwd_kernel, dim3((rows + rpc - 1) / rpc), dim3(256), 0, static_cast<cudaStream_t>(stream_), p));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_colsum_bf16(const void* G, long long ld, int rows, int C, float* out, void* stream_);

extern "C" int hma_cast_transpose(const float* W, int R, int Cc, void* Wb, void* Wt, float alpha, void* stream_) {
  using namespace hma;
  if (R == 0 || Cc == 0) return 0;
  dim3 grid((Cc + 31) / 32, (R + 31) / 32);
  cast_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      W, R, Cc, static_cast<__nv_bfloat16*>(Wb), static_cast<__nv_bfloat16*>(Wt), alpha);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace hma {

__global__ void __launch_bounds__(256) cast_flat_kernel(const float* x, __nv_bfloat16* y, long long count4) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

// ActionStat (st_mask_git.py:134-138): (a - mean[j % adim]) / (std[j % adim] + 1e-10), written as a
// zero-padded bf16 [rows, kpad] operand for the stem's first projection.
__global__ void __launch_bounds__(256) action_prep_kernel(const float* a, int rows, int da, const float* mean,
                                                         const float* stdv, int adim, __nv_bfloat16* y, int kpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * kpad) return;
  const int j = (int)(i % kpad);
  const long long r = i / kpad;
  float v = 0.f;
  if (j < da) {
    v = a[r * da + j];
    if (mean != nullptr) v = (v - mean[j % adim]) / (stdv[j % adim] + 1e-10f);
  }
  y[i] = __float2bfloat16(v);
}

// relu(LayerNorm(x)) for the stem (st_mask_git.py:94-96) and its backward, fp32 rows of 256.
__global__ void __launch_bounds__(256) ln_relu_fwd_kernel(const float* x, int rows, const float* gamma,
                                                         const float* beta, float eps, __nv_bfloat16* y,
                                                         float* stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const RowLoad r = load_row_f32(x + (size_t)row * kC, lane);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += r.v[j];
  const float mean = warp_sum(s) * (1.0f / kC);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = r.v[j] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kC) + eps);
  if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  float out[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col_of(lane, j);
    out[j] = fmaxf((r.v[j] - mean) * rstd * gamma[c] + beta[c], 0.f);
  }
  store_row_bf16(y + (size_t)row * kC, lane, out);
}

// dx (fp32, written) from dy (fp32, grad wrt relu output); dgamma / dbeta accumulated with atomics.
__global__ void __launch_bounds__(256) ln_relu_bwd_kernel(const float* dy, const float* x, const float* stats, int rows,
                                                         const float* gamma, const float* beta, float* dx,
                                                         float* dgamma, float* dbeta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const RowLoad xr = load_row_f32(x + (size_t)row * kC, lane);
  const RowLoad g0 = load_row_f32(dy + (size_t)row * kC, lane);
  const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
  float xh[8], g[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col_of(lane, j);
    xh[j] = (xr.v[j] - mean) * rstd;
    const float pre = xh[j] * gamma[c] + beta[c];
    const float d = pre > 0.f ? g0.v[j] : 0.f;
    atomicAdd(dgamma + c, d * xh[j]);
    atomicAdd(dbeta + c, d);
    g[j] = d * gamma[c];
    s1 += g[j];
    s2 = fmaf(g[j], xh[j], s2);
  }
  s1 = warp_sum(s1) * (1.0f / kC);
  s2 = warp_sum(s2) * (1.0f / kC);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] - s1 - xh[j] * s2);
  store_row_f32(dx + (size_t)row * kC, lane, o);
}

// out[c] += sum_r G[r, c] for fp32 G (small matrices: stem / adaLN bias gradients)
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* G, long long ld, int rows, int C, float* out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int r = 0; r < rows; ++r) a += G[(size_t)r * ld + c];
  out[c] += a;
}

}  // namespace hma

extern "C" int hma_cast_bf16(const float* x, void* y, long long count, void* stream_) {
  using namespace hma;
  if (count == 0) return 0;
  HMA_REQUIRE(count % 4 == 0, "cast_bf16: element count must be a multiple of 4");
  const long long c4 = count / 4;
  HMA_CHECK_CUDA(hma_host::launch_pdl(cast_flat_kernel, dim3((unsigned)((c4 + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, static_cast<__nv_bfloat16*>(y), c4));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_action_prep(const float* a, int rows, int da, const float* mean, const float* stdv, int adim,
                               void* y, int kpad, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(kpad >= da && kpad % 64 == 0, "action_prep: kpad=%d must be a multiple of 64 >= d_action=%d", kpad, da);
  HMA_REQUIRE(mean == nullptr || (adim > 0 && da % adim == 0), "action_prep: d_action must be a multiple of action_dim");
  const long long total = (long long)rows * kpad;
  action_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      a, rows, da, mean, stdv, adim, static_cast<__nv_bfloat16*>(y), kpad);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_ln_relu_fwd(const float* x, int rows, const float* gamma, const float* beta, float eps, void* y,
                               float* stats, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  ln_relu_fwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, rows, gamma, beta, eps, static_cast<__nv_bfloat16*>(y), stats);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_ln_relu_bwd(const float* dy, const float* x, const float* stats, int rows, const float* gamma,
                               const float* beta, float* dx, float* dgamma, float* dbeta, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  ln_relu_bwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(dy, x, stats, rows, gamma, beta,
                                                                                   dx, dgamma, dbeta);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_colsum_f32(const float* G, long long ld, int rows, int C, float* out, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  colsum_f32_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(G, ld, rows, C, out);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace hma {
// dst[(f*n + s), :] = s < S ? src[(f*S + s), :] : 0   (video-token rows back into the full token grid)
__global__ void __launch_bounds__(256) rows_scatter_kernel(const float* src, float* dst, int frames, int S, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)frames * n) return;
  const int s = (int)(row % n);
  const long long f = row / n;
  float4 a = make_float4(0, 0, 0, 0), b = a;
  if (s < S) {
    const float* sr = src + ((size_t)f * S + s) * kC;
    a = *reinterpret_cast<const float4*>(sr + lane * 4);
    b = *reinterpret_cast<const float4*>(sr + 128 + lane * 4);
  }
  float* d = dst + (size_t)row * kC;
  *reinterpret_cast<float4*>(d + lane * 4) = a;
  *reinterpret_cast<float4*>(d + 128 + lane * 4) = b;
}
}  // namespace hma

extern "C" int hma_rows_scatter(const float* src, float* dst, int frames, int S, int n, void* stream_) {
  using namespace hma;
  const long long rows = (long long)frames * n;
  if (rows == 0) return 0;
  HMA_REQUIRE(S <= n, "rows_scatter: S=%d > n=%d", S, n);
  rows_scatter_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, dst, frames, S, n);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace hma {

// y = bf16(x) for fp32 rows of 256, plus (optionally) colsum[c] += sum_r x[r, c] of the ROUNDED values:
// the bias gradient of the projection that consumes y, for free while the row streams by.
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* x, __nv_bfloat16* y, int rows, float* colsum) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 64;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int i = warp; i < 64; i += 8) {
    const int row = r0 + i;
    if (row < rows) {
      const RowLoad r = load_row_f32(x + (size_t)row * kC, lane);
      const uint2 a = make_uint2(pack_bf16(r.v[0], r.v[1]), pack_bf16(r.v[2], r.v[3]));
      const uint2 b = make_uint2(pack_bf16(r.v[4], r.v[5]), pack_bf16(r.v[6], r.v[7]));
      *reinterpret_cast<uint2*>(y + (size_t)row * kC + lane * 4) = a;
      *reinterpret_cast<uint2*>(y + (size_t)row * kC + 128 + lane * 4) = b;
      acc[0] += bf16_lo(a.x); acc[1] += bf16_hi(a.x); acc[2] += bf16_lo(a.y); acc[3] += bf16_hi(a.y);
      acc[4] += bf16_lo(b.x); acc[5] += bf16_hi(b.x); acc[6] += bf16_lo(b.y); acc[7] += bf16_hi(b.y);
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][col_of(lane, j)] = acc[j];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
  atomicAdd(colsum + threadIdx.x, s);
}

// out[c] += sum_r G[r, c], bf16 G with C % 8 == 0, C <= 2048: 16-byte loads, 32-row slabs per CTA (a thread walks at most
// 16 rows, all loads in flight at once; 128-row slabs were latency-bound at 34 us for a 19 MB matrix).
constexpr int kColsumSlab = 32;
__global__ void __launch_bounds__(256) colsum_wide_kernel(const __nv_bfloat16* G, long long ld, int rows, int C,
                                                          float* out) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float wred[];  // [groups][C]
  const int vec_per_row = C / 8;
  const int groups = 256 / vec_per_row;  // row groups processed concurrently (>= 1)
  const int g = threadIdx.x / vec_per_row, vcol = threadIdx.x % vec_per_row;
  const int r0 = blockIdx.x * kColsumSlab;
  const int r1 = min(rows, r0 + kColsumSlab);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (g < groups) {
#pragma unroll 8
    for (int r = r0 + g; r < r1; r += groups) {
      const uint4 v = *reinterpret_cast<const uint4*>(G + (size_t)r * ld + vcol * 8);
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) wred[g * C + vcol * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (int q = 0; q < groups; ++q) s += wred[q * C + c];
    atomicAdd(out + c, s);
  }
}

struct CastDesc {
  const float* src;
  __nv_bfloat16* plain;
  __nv_bfloat16* trans;
  long long R, C;
};

// One launch for every weight matrix of a step: blockIdx.z selects the matrix.
__global__ void __launch_bounds__(256) cast_transpose_batched_kernel(const CastDesc* descs) {
  __shared__ float tile[32][33];
  const CastDesc d = descs[blockIdx.z];
  const int R = (int)d.R, Cc = (int)d.C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (c0 >= Cc || r0 >= R) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = d.src[(size_t)r * Cc + c];
      if (d.plain != nullptr) d.plain[(size_t)r * Cc + c] = __float2bfloat16(v);
    }
    tile[i][tx] = v;
  }
  if (d.trans == nullptr) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < Cc) d.trans[(size_t)c * R + r] = __float2bfloat16(tile[tx][i]);
  }
}

}  // namespace hma

extern "C" int hma_cast_colsum(const float* x, void* y, int rows, float* colsum, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_CHECK_CUDA(hma_host::launch_pdl(cast_colsum_kernel, dim3((rows + 63) / 64), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, static_cast<__nv_bfloat16*>(y), rows, colsum));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_cast_transpose_batched(const void* descs, int count, int max_rows, int max_cols, void* stream_) {
  using namespace hma;
  if (count == 0) return 0;
  HMA_REQUIRE(count <= 65535, "cast_transpose_batched: too many matrices");
  dim3 grid((max_cols + 31) / 32, (max_rows + 31) / 32, count);
  cast_transpose_batched_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const CastDesc*>(descs));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_colsum_bf16(const void* G, long long ld, int rows, int C, float* out, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(C % 8 == 0 && C <= 2048 && ld % 8 == 0, "colsum: C=%d must be a multiple of 8, <= 2048", C);
  const int vec_per_row = C / 8;
  const int groups = 256 / vec_per_row;
  HMA_REQUIRE(groups >= 1, "colsum: C too wide");
  const size_t smem = (size_t)groups * C * sizeof(float);
  HMA_CHECK_CUDA(hma_host::launch_pdl(colsum_wide_kernel, dim3((rows + kColsumSlab - 1) / kColsumSlab), dim3(256), smem, static_cast<cudaStream_t>(stream_), static_cast<const __nv_bfloat16*>(G), ld, rows, C, out));
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Additive action conditioning (action_network containing "mlp", st_transformer.py:93-97): the per-(b, t) action
// embedding is added to every token of its frame in every layer. Forward: y[r] = x[r] + v[r / rows_per_group];
// backward: dv[g] += sum of the group's rows of dx (dx itself passes through unchanged).
// ------------------------------------------------------------------------------------------------
namespace hma {
__global__ void __launch_bounds__(256) group_add_kernel(const float* x, const float* v, float* y, int rows, int rows_per_group) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    const RowLoad a = load_row_f32(x + row * kC, lane);
    const RowLoad b = load_row_f32(v + (row / rows_per_group) * kC, lane);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a.v[j] + b.v[j];
    store_row_f32(y + row * kC, lane, o);
  }
}
__global__ void __launch_bounds__(256) group_colsum_kernel(const float* dx, float* dv, int rows_per_group) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][kC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = dx + (size_t)blockIdx.x * rows_per_group * kC;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int r = warp; r < rows_per_group; r += 8) {
    const RowLoad a = load_row_f32(base + (size_t)r * kC, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += a.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][col_of(lane, j)] = acc[j];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
  dv[(size_t)blockIdx.x * kC + threadIdx.x] += s;  // one CTA per group: no atomics
}
}  // namespace hma

extern "C" int hma_group_add(const float* x, const float* v, float* y, int rows, int rows_per_group, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(rows_per_group > 0 && rows % rows_per_group == 0, "group_add: rows=%d is not a multiple of rows_per_group=%d", rows, rows_per_group);
  int grid = (rows + 7) / 8;
  const int cap = hma_host::sm_count() * 8;
  if (grid > cap) grid = cap;
  HMA_CHECK_CUDA(hma_host::launch_pdl(group_add_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, v, y, rows, rows_per_group));
  return 0;
}

extern "C" int hma_group_colsum(const float* dx, float* dv, int groups, int rows_per_group, void* stream_) {
  using namespace hma;
  if (groups == 0) return 0;
  HMA_REQUIRE(rows_per_group > 0, "group_colsum: bad rows_per_group");
  HMA_CHECK_CUDA(hma_host::launch_pdl(group_colsum_kernel, dim3(groups), dim3(256), 0, static_cast<cudaStream_t>(stream_), dx, dv, rows_per_group));
  return 0;
}
