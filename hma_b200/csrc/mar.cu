// STMAR (continuous-token) stages around the shared ST trunk and the tensor-core GEMMs
// (SURVEY.md §8a rows R1-R3). Everything here is HBM-bound row-wise work:
//   mar_embed_fwd/bwd   : mask-token fill + patchify + Linear(D->256) + action-token concat + positional embedding
//                         (st_mar.py:146-176,199-207,240)
//   mar_ln_fwd/bwd      : LayerNorm over C in {256, 1024} with optional affine, per-row adaLN shift/scale read from a bf16
//                         modulation matrix, and an additive row table — z_proj_ln, decoder_norm + diffusion_pos_embed
//                         (st_mar.py:174,190-191), ResBlock.in_ln + modulate, FinalLayer (diffloss.py:116-159)
//   mar_gate_fwd/bwd    : x + gate * h (diffloss.py:140)
//   mar_silu_fwd/bwd    : SiLU of the conditioning vector y = t_emb + c_emb (diffloss.py:127,150)
//   mar_q_sample, mar_timestep_embed : q(x_t | x_0) and the sinusoidal embedding (gaussian_diffusion.py:200-215; diffloss.py:80-100)
//   mar_diff_loss_*     : MSE + variational-bound loss with learned-range variance and its gradient
//                         (gaussian_diffusion.py:650-745; diffusion_utils.py:10-64; diffloss.py:28-35)
//   mar_p_sample        : one ancestral DDPM step (gaussian_diffusion.py:237-314,358-392)
//   mar_gather/scatter_rows, dropout_* : MaskGIT token bookkeeping (st_mar.py:414-446) and nn.Dropout (st_transformer.py:24-27)
// One warp per row for the LayerNorm kernels (a lane owns fixed columns, so column reductions stay in registers across a
// grid-stride loop); grids sized to a multiple of the SM count.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

static inline int grid_for(long long work_items, int per_block, int max_per_sm = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)hma_host::sm_count() * max_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------------------------------------------
// front end
// ---------------------------------------------------------------------------------------------------------------
struct EmbedDims {
  int B, T, H, W, Cv, p, A, pos_n;
  int hp, wp, Sp, D, n;  // derived
};

__device__ __forceinline__ long long pixel_index(const EmbedDims& d, int bt, int s, int e, int* cv) {
  // patch slot s = (hy, wx); element e = (pi*p + qi)*Cv + c  (einsum "nthpwqc->nthwpqc", st_mar.py:199-207)
  const int hy = s / d.wp, wx = s % d.wp;
  const int c = e % d.Cv, pq = e / d.Cv;
  const int pi = pq / d.p, qi = pq % d.p;
  *cv = c;
  return (((long long)bt * d.H + hy * d.p + pi) * d.W + wx * d.p + qi);
}

__global__ void __launch_bounds__(256) mar_embed_fwd_kernel(float* lat, const unsigned char* mask, const float* mask_token,
                                                            const float* xp_in, const float* We, const float* act,
                                                            const float* pos, EmbedDims d, int fill_inplace, float* u,
                                                            float* xp_out, float* rowmask) {
  extern __shared__ float sW[];  // [D][256] transposed weight
  pdl_wait();
  for (int i = threadIdx.x; i < d.D * 256; i += blockDim.x) {
    const int e = i / 256, c = i % 256;
    sW[i] = We[c * d.D + e];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)d.B * d.T * d.n;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const int bt = (int)(r / d.n), s = (int)(r % d.n), t = bt % d.T;
    const float* prow = pos + ((long long)t * d.pos_n + s) * 256;
    float* urow = u + r * 256;
    if (s >= d.Sp) {  // action token: the frame's action embedding, replicated (st_mar.py:164-170)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        urow[c] = act[(long long)bt * 256 + c] + prow[c];
      }
      continue;
    }
    float v0 = 0.f, v1 = 0.f;  // this lane's patch elements e = lane, lane + 32
    bool any_masked = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = lane + 32 * h;
      if (e < d.D) {
        float v;
        if (xp_in != nullptr) {
          v = xp_in[((long long)bt * d.Sp + s) * d.D + e];
        } else {
          int cv;
          const long long pix = pixel_index(d, bt, s, e, &cv);
          if (mask != nullptr && mask[pix]) {
            any_masked = true;
            v = mask_token[cv];
            if (fill_inplace) lat[pix * d.Cv + cv] = v;
          } else {
            v = lat[pix * d.Cv + cv];
          }
        }
        if (xp_out != nullptr) xp_out[((long long)bt * d.Sp + s) * d.D + e] = v;
        if (h == 0) v0 = v; else v1 = v;
      }
    }
    any_masked = __any_sync(0xffffffffu, any_masked);  // st_mar.py:252: a patch is relevant if any of its pixels is masked
    if (rowmask != nullptr && lane == 0) rowmask[(long long)bt * d.Sp + s] = any_masked ? 1.f : 0.f;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = prow[lane + 32 * j];
    for (int e = 0; e < d.D; ++e) {
      const float xe = __shfl_sync(0xffffffffu, e < 32 ? v0 : v1, e & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(xe, sW[e * 256 + lane + 32 * j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) urow[lane + 32 * j] = acc[j];
  }
}

__global__ void __launch_bounds__(256) mar_embed_bwd_kernel(const float* du, const float* xp, const unsigned char* mask,
                                                            const float* We, EmbedDims d, float* dWe, float* dmask_token,
                                                            float* dact, float* dpos) {
  extern __shared__ float sm[];
  float* sW = sm;                 // [D][256]
  float* sdW = sm + d.D * 256;    // [D][256] partial dWe
  float* sdm = sdW + d.D * 256;   // [Cv] partial dmask_token
  pdl_wait();
  for (int i = threadIdx.x; i < d.D * 256; i += blockDim.x) {
    sW[i] = We[(i % 256) * d.D + i / 256];
    sdW[i] = 0.f;
  }
  if (threadIdx.x < d.Cv) sdm[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)d.B * d.T * d.n;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const int bt = (int)(r / d.n), s = (int)(r % d.n), t = bt % d.T;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = du[r * 256 + lane + 32 * j];
    float* dprow = dpos + ((long long)t * d.pos_n + s) * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dprow + lane + 32 * j, g[j]);
    if (s >= d.Sp) {
      if (dact != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dact + (long long)bt * 256 + lane + 32 * j, g[j]);
      }
      continue;
    }
    const float* xrow = xp + ((long long)bt * d.Sp + s) * d.D;
    for (int e = 0; e < d.D; ++e) {
      const float xe = xrow[e];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(sdW + e * 256 + lane + 32 * j, g[j] * xe);
        dot = fmaf(g[j], sW[e * 256 + lane + 32 * j], dot);
      }
      if (mask != nullptr && dmask_token != nullptr) {  // d(patch element) reaches mask_token where the pixel was masked
        dot = warp_sum(dot);
        if (lane == 0) {
          int cv;
          const long long pix = pixel_index(d, bt, s, e, &cv);
          if (mask[pix]) atomicAdd(sdm + cv, dot);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.D * 256; i += blockDim.x) atomicAdd(dWe + (i % 256) * d.D + i / 256, sdW[i]);
  if (threadIdx.x < d.Cv && dmask_token != nullptr) atomicAdd(dmask_token + threadIdx.x, sdm[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm with optional affine / modulation / additive table; C = 128 * V
// ---------------------------------------------------------------------------------------------------------------
struct LnArgs {
  const float* x;
  int rows;
  const float* gamma;
  const float* beta;
  float eps;
  const __nv_bfloat16* mod;
  long long ldmod;
  int shift_off, scale_off;
  const float* add;
  int add_rows;
  float* y32;
  __nv_bfloat16* y16;
  float* stats;
  // optional fused residual gate of the PREVIOUS block (diffloss.py:140): the row normalised is x + gmod[gate_off..] * h2,
  // which is also written to xsum (the new residual stream)
  const __nv_bfloat16* gmod;
  long long ldg;
  int gate_off;
  const __nv_bfloat16* h2;
  float* xsum;
};

template <int V>
__global__ void __launch_bounds__(256) mar_ln_fwd_kernel(const LnArgs a) {
  constexpr int C = 128 * V;
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < a.rows; r += (long long)gridDim.x * 8) {
    float4 v[V];
    uint2 sh[V], sc[V];  // modulation of this row, requested before the statistics so that its latency overlaps them
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k] = reinterpret_cast<const float4*>(a.x + r * C)[k * 32 + lane];
      if (a.mod != nullptr) {
        const int col = (k * 32 + lane) * 4;
        sh[k] = *reinterpret_cast<const uint2*>(a.mod + r * a.ldmod + a.shift_off + col);
        sc[k] = *reinterpret_cast<const uint2*>(a.mod + r * a.ldmod + a.scale_off + col);
      }
    }
    if (a.h2 != nullptr) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const int col = (k * 32 + lane) * 4;
        const uint2 g = *reinterpret_cast<const uint2*>(a.gmod + r * a.ldg + a.gate_off + col);
        const uint2 h = *reinterpret_cast<const uint2*>(a.h2 + r * C + col);
        v[k].x = fmaf(bf16_lo(g.x), bf16_lo(h.x), v[k].x);
        v[k].y = fmaf(bf16_hi(g.x), bf16_hi(h.x), v[k].y);
        v[k].z = fmaf(bf16_lo(g.y), bf16_lo(h.y), v[k].z);
        v[k].w = fmaf(bf16_hi(g.y), bf16_hi(h.y), v[k].w);
        if (a.xsum != nullptr) *reinterpret_cast<float4*>(a.xsum + r * C + col) = v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
      q += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + a.eps);
    if (a.stats != nullptr && lane == 0) {
      a.stats[r * 2] = mean;
      a.stats[r * 2 + 1] = rstd;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int col = (k * 32 + lane) * 4;
      float o[4] = {v[k].x * rstd, v[k].y * rstd, v[k].z * rstd, v[k].w * rstd};
      if (a.gamma != nullptr) {
        const float4 gm = *reinterpret_cast<const float4*>(a.gamma + col);
        const float4 bt = *reinterpret_cast<const float4*>(a.beta + col);
        o[0] = o[0] * gm.x + bt.x; o[1] = o[1] * gm.y + bt.y; o[2] = o[2] * gm.z + bt.z; o[3] = o[3] * gm.w + bt.w;
      }
      if (a.mod != nullptr) {
        o[0] = o[0] * (1.f + bf16_lo(sc[k].x)) + bf16_lo(sh[k].x);
        o[1] = o[1] * (1.f + bf16_hi(sc[k].x)) + bf16_hi(sh[k].x);
        o[2] = o[2] * (1.f + bf16_lo(sc[k].y)) + bf16_lo(sh[k].y);
        o[3] = o[3] * (1.f + bf16_hi(sc[k].y)) + bf16_hi(sh[k].y);
      }
      if (a.add != nullptr) {
        const float4 ad = *reinterpret_cast<const float4*>(a.add + (r % a.add_rows) * C + col);
        o[0] += ad.x; o[1] += ad.y; o[2] += ad.z; o[3] += ad.w;
      }
      if (a.y32 != nullptr) *reinterpret_cast<float4*>(a.y32 + r * C + col) = make_float4(o[0], o[1], o[2], o[3]);
      if (a.y16 != nullptr)
        *reinterpret_cast<uint2*>(a.y16 + r * C + col) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
  }
}

struct LnBwdArgs {
  const __nv_bfloat16* dy16;
  const float* dy32;
  const float* x;
  const float* stats;
  int rows;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* mod;
  long long ldmod;
  int shift_off, scale_off;
  float* dx32;
  int accumulate;
  __nv_bfloat16* dx16;
  float* dgamma;
  float* dbeta;
  __nv_bfloat16* dmod;
  long long lddmod;
  float* dadd;
  int add_rows;
};

template <int V>
__global__ void __launch_bounds__(256) mar_ln_bwd_kernel(const LnBwdArgs a) {
  constexpr int C = 128 * V;
  __shared__ float red[8][128];
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dg[V][4], db[V][4];
#pragma unroll
  for (int k = 0; k < V; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[k][e] = db[k][e] = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < a.rows; r += (long long)gridDim.x * 8) {
    // Phase 1: every load of the row is issued up front through the read-only path. (With plain loads interleaved with
    // the dmod / dx stores below the compiler had to keep program order — possible aliasing — which serialised eight
    // DRAM round trips per row: 115 us for 6144 rows of 1024.)
    const float mean = __ldg(a.stats + r * 2), rstd = __ldg(a.stats + r * 2 + 1);
    float4 xv[V], gv[V];
    uint2 sc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int col = (k * 32 + lane) * 4;
      xv[k] = __ldg(reinterpret_cast<const float4*>(a.x + r * C + col));
      if (a.dy16 != nullptr) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(a.dy16 + r * C + col));
        gv[k] = make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
      } else {
        gv[k] = __ldg(reinterpret_cast<const float4*>(a.dy32 + r * C + col));
      }
      if (a.mod != nullptr) sc[k] = __ldg(reinterpret_cast<const uint2*>(a.mod + r * a.ldmod + a.scale_off + col));
    }
    float4 old[V];
    if (a.dx32 != nullptr && a.accumulate) {
#pragma unroll
      for (int k = 0; k < V; ++k) old[k] = *reinterpret_cast<const float4*>(a.dx32 + r * C + (k * 32 + lane) * 4);
    }
    // Phase 2: arithmetic + the modulation-gradient stores
    float dxn[V][4], xn[V][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int col = (k * 32 + lane) * 4;
      const float g[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
      xn[k][0] = (xv[k].x - mean) * rstd; xn[k][1] = (xv[k].y - mean) * rstd;
      xn[k][2] = (xv[k].z - mean) * rstd; xn[k][3] = (xv[k].w - mean) * rstd;
      if (a.dadd != nullptr) {
        float* dst = a.dadd + (r % a.add_rows) * C + col;
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(dst + e, g[e]);
      }
      float gm[4] = {1.f, 1.f, 1.f, 1.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.gamma != nullptr) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + col));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta + col));
        gm[0] = g4.x; gm[1] = g4.y; gm[2] = g4.z; gm[3] = g4.w;
        bt[0] = b4.x; bt[1] = b4.y; bt[2] = b4.z; bt[3] = b4.w;
      }
      float da[4] = {g[0], g[1], g[2], g[3]};
      if (a.mod != nullptr) {
        const float scl[4] = {bf16_lo(sc[k].x), bf16_hi(sc[k].x), bf16_lo(sc[k].y), bf16_hi(sc[k].y)};
        float dsc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dsc[e] = g[e] * (xn[k][e] * gm[e] + bt[e]);
          da[e] = g[e] * (1.f + scl[e]);
        }
        if (a.dmod != nullptr) {
          *reinterpret_cast<uint2*>(a.dmod + r * a.lddmod + a.shift_off + col) =
              make_uint2(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]));
          *reinterpret_cast<uint2*>(a.dmod + r * a.lddmod + a.scale_off + col) =
              make_uint2(pack_bf16(dsc[0], dsc[1]), pack_bf16(dsc[2], dsc[3]));
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dg[k][e] += da[e] * xn[k][e];
        db[k][e] += da[e];
        dxn[k][e] = da[e] * gm[e];
        s1 += dxn[k][e];
        s2 += dxn[k][e] * xn[k][e];
      }
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int col = (k * 32 + lane) * 4;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = rstd * (dxn[k][e] - s1 - xn[k][e] * s2);
      if (a.dx32 != nullptr) {
        if (a.accumulate) {
          o[0] += old[k].x; o[1] += old[k].y; o[2] += old[k].z; o[3] += old[k].w;
        }
        *reinterpret_cast<float4*>(a.dx32 + r * C + col) = make_float4(o[0], o[1], o[2], o[3]);
      }
      if (a.dx16 != nullptr)
        *reinterpret_cast<uint2*>(a.dx16 + r * C + col) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
  }
  if (a.dgamma == nullptr) return;
  // column sums across the block's 8 warps, 128 columns at a time, then one atomic per column per block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? a.dgamma : a.dbeta;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) red[warp][lane * 4 + e] = pass == 0 ? dg[k][e] : db[k][e];
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(dst + k * 128 + threadIdx.x, s);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise pieces of the diffusion MLP
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mar_gate_fwd_kernel(const float* x, const __nv_bfloat16* mod, long long ldmod,
                                                           int gate_off, const __nv_bfloat16* h2, long long rows, int C,
                                                           float* out) {
  pdl_wait();
  const long long total = rows * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C / 4);
    const int col = (int)(i % (C / 4)) * 4;
    const float4 xv = *reinterpret_cast<const float4*>(x + r * C + col);
    const uint2 g = *reinterpret_cast<const uint2*>(mod + r * ldmod + gate_off + col);
    const uint2 h = *reinterpret_cast<const uint2*>(h2 + r * C + col);
    float4 o;
    o.x = fmaf(bf16_lo(g.x), bf16_lo(h.x), xv.x);
    o.y = fmaf(bf16_hi(g.x), bf16_hi(h.x), xv.y);
    o.z = fmaf(bf16_lo(g.y), bf16_lo(h.y), xv.z);
    o.w = fmaf(bf16_hi(g.y), bf16_hi(h.y), xv.w);
    *reinterpret_cast<float4*>(out + r * C + col) = o;
  }
}

__global__ void __launch_bounds__(256) mar_gate_bwd_kernel(const float* dx, const __nv_bfloat16* mod, long long ldmod,
                                                           int gate_off, const __nv_bfloat16* h2, long long rows, int C,
                                                           __nv_bfloat16* dh2, __nv_bfloat16* dmod, long long lddmod) {
  pdl_wait();
  const long long total = rows * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C / 4);
    const int col = (int)(i % (C / 4)) * 4;
    const float4 d = *reinterpret_cast<const float4*>(dx + r * C + col);
    const uint2 g = *reinterpret_cast<const uint2*>(mod + r * ldmod + gate_off + col);
    const uint2 h = *reinterpret_cast<const uint2*>(h2 + r * C + col);
    *reinterpret_cast<uint2*>(dh2 + r * C + col) =
        make_uint2(pack_bf16(d.x * bf16_lo(g.x), d.y * bf16_hi(g.x)), pack_bf16(d.z * bf16_lo(g.y), d.w * bf16_hi(g.y)));
    *reinterpret_cast<uint2*>(dmod + r * lddmod + gate_off + col) =
        make_uint2(pack_bf16(d.x * bf16_lo(h.x), d.y * bf16_hi(h.x)), pack_bf16(d.z * bf16_lo(h.y), d.w * bf16_hi(h.y)));
  }
}

__device__ __forceinline__ float silu_exact(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float dsilu_exact(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}

__global__ void __launch_bounds__(256) mar_silu_fwd_kernel(const float* y, const float* rowvec, long long rows, int C,
                                                           __nv_bfloat16* out) {
  pdl_wait();
  const long long total = rows * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (C / 4)) * 4;
    float4 v = reinterpret_cast<const float4*>(y)[i];
    if (rowvec != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(rowvec + col);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    reinterpret_cast<uint2*>(out)[i] =
        make_uint2(pack_bf16(silu_exact(v.x), silu_exact(v.y)), pack_bf16(silu_exact(v.z), silu_exact(v.w)));
  }
}

// out[(i*n + r), :] = SiLU(c[r, :] + te[i, :]) for every spaced step i: the conditioning of ALL sampler steps at once
__global__ void __launch_bounds__(256) mar_silu_steps_kernel(const float* c, const float* te, long long n, int steps, int C,
                                                            __nv_bfloat16* out) {
  pdl_wait();
  const long long per = n * (C / 4);
  const long long total = per * steps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rem = i % per;
    const int col = (int)(rem % (C / 4)) * 4;
    const float4 v = reinterpret_cast<const float4*>(c)[rem];
    const float4 a = *reinterpret_cast<const float4*>(te + (i / per) * C + col);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(silu_exact(v.x + a.x), silu_exact(v.y + a.y)),
                                                 pack_bf16(silu_exact(v.z + a.z), silu_exact(v.w + a.w)));
  }
}

__global__ void __launch_bounds__(256) mar_silu_bwd_kernel(const float* dsy, const float* y, long long n4, __nv_bfloat16* dy) {
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(dsy)[i];
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    reinterpret_cast<uint2*>(dy)[i] = make_uint2(pack_bf16(g.x * dsilu_exact(v.x), g.y * dsilu_exact(v.y)),
                                                 pack_bf16(g.z * dsilu_exact(v.z), g.w * dsilu_exact(v.w)));
  }
}

// tables: fp32 [steps, 8] = sqrt_acp, sqrt_1m_acp, sqrt_recip_acp, sqrt_recipm1_acp, coef1, coef2, post_logvar, log_beta
__global__ void __launch_bounds__(256) mar_q_sample_kernel(const float* x0, const float* noise, const long long* t,
                                                           const float* tables, long long N, int D, int kpad,
                                                           __nv_bfloat16* xt16) {
  pdl_wait();
  const long long total = N * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kpad;
    const int e = (int)(i % kpad);
    float v = 0.f;
    if (e < D) {
      v = x0[r * D + e];
      if (noise != nullptr) {
        const float* tb = tables + t[r] * 8;
        v = tb[0] * v + tb[1] * noise[r * D + e];
      }
    }
    xt16[i] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(256) mar_timestep_embed_kernel(const long long* t, long long N, __nv_bfloat16* out) {
  pdl_wait();
  const long long total = N * 256;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 256;
    const int j = (int)(i % 256);
    const float freq = expf(-9.210340371976184f * (float)(j & 127) * (1.f / 128.f));  // diffloss.py:91-93
    const float arg = (float)t[r] * freq;
    out[i] = __float2bfloat16(j < 128 ? cosf(arg) : sinf(arg));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// diffusion loss
// ---------------------------------------------------------------------------------------------------------------
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kCdfA = 0.7978845608028654f;  // sqrt(2/pi)

__device__ __forceinline__ float approx_cdf(float u) { return 0.5f * (1.f + tanhf(kCdfA * (u + 0.044715f * u * u * u))); }
__device__ __forceinline__ float approx_pdf(float u) {  // d approx_cdf / du
  const float th = tanhf(kCdfA * (u + 0.044715f * u * u * u));
  return 0.5f * (1.f - th * th) * kCdfA * (1.f + 3.f * 0.044715f * u * u);
}

// One element's vb term and its derivative w.r.t. the model log-variance (the mean prediction is detached in the vb
// term, gaussian_diffusion.py:703-712).
__device__ __forceinline__ void vb_element(float x0, float xt, float eps_hat, float v, const float* tb, bool t0, float* term,
                                           float* dterm_dv) {
  const float tlv = tb[6], lb = tb[7];
  const float frac = 0.5f * (v + 1.f);
  const float lv = frac * lb + (1.f - frac) * tlv;
  const float dlv_dv = 0.5f * (lb - tlv);
  const float px0 = tb[2] * xt - tb[3] * eps_hat;
  const float mean = tb[4] * px0 + tb[5] * xt;
  if (!t0) {
    const float tm = tb[4] * x0 + tb[5] * xt;
    const float e1 = expf(tlv - lv), e2 = expf(-lv), dm = (tm - mean) * (tm - mean);
    *term = 0.5f * (-1.f + lv - tlv + e1 + dm * e2);
    *dterm_dv = 0.5f * (1.f - e1 - dm * e2) * dlv_dv;
    return;
  }
  // decoder NLL: -log of the discretised Gaussian likelihood (diffusion_utils.py:38-64), log_scales = lv / 2
  const float inv = expf(-0.5f * lv), cx = x0 - mean;
  const float up = inv * (cx + 1.f / 255.f), um = inv * (cx - 1.f / 255.f);
  const float cp = approx_cdf(up), cm = approx_cdf(um);
  float logp, dlogp_dls;  // d u / d log_scale = -u
  if (x0 < -0.999f) {
    logp = logf(fmaxf(cp, 1e-12f));
    dlogp_dls = cp >= 1e-12f ? approx_pdf(up) * (-up) / cp : 0.f;
  } else if (x0 > 0.999f) {
    const float q = 1.f - cm;
    logp = logf(fmaxf(q, 1e-12f));
    dlogp_dls = q >= 1e-12f ? approx_pdf(um) * um / q : 0.f;
  } else {
    const float dl = cp - cm;
    logp = logf(fmaxf(dl, 1e-12f));
    dlogp_dls = dl >= 1e-12f ? (approx_pdf(up) * (-up) + approx_pdf(um) * um) / dl : 0.f;
  }
  *term = -logp;
  *dterm_dv = -dlogp_dls * 0.5f * dlv_dv;
}

__global__ void __launch_bounds__(128) mar_diff_loss_fwd_kernel(const float* out, long long ldo, const float* x0,
                                                                const float* noise, const long long* t, const float* mask,
                                                                const float* tables, long long N, int D, float* rows_loss,
                                                                float* sums) {
  pdl_wait();
  float s_loss = 0.f, s_mask = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
    const long long tr = t[r];
    const float* tb = tables + tr * 8;
    float mse = 0.f, vb = 0.f;
    for (int e = 0; e < D; ++e) {
      const float a = x0[r * D + e], nz = noise[r * D + e];
      const float xt = tb[0] * a + tb[1] * nz;
      const float eh = out[r * ldo + e], v = out[r * ldo + D + e];
      mse += (nz - eh) * (nz - eh);
      float term, dv;
      vb_element(a, xt, eh, v, tb, tr == 0, &term, &dv);
      vb += term;
    }
    const float row = mse / D + vb / (D * kLn2);
    if (rows_loss != nullptr) rows_loss[r] = row;
    const float m = mask != nullptr ? mask[r] : 1.f;
    s_loss += row * m;
    s_mask += m;
  }
  s_loss = warp_sum(s_loss);
  s_mask = warp_sum(s_mask);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums, s_loss);
    atomicAdd(sums + 1, s_mask);
  }
}

__global__ void mar_diff_loss_finish_kernel(const float* sums, long long N, int has_mask, float* loss) {
  pdl_wait();
  *loss = has_mask ? sums[0] / (sums[1] + 1e-8f) : sums[0] / (float)N;  // diffloss.py:33-35
}

__global__ void __launch_bounds__(128) mar_diff_loss_bwd_kernel(const float* out, long long ldo, const float* x0,
                                                                const float* noise, const long long* t, const float* mask,
                                                                const float* tables, long long N, int D, const float* sums,
                                                                const float* dloss, __nv_bfloat16* dout, long long ldd) {
  pdl_wait();
  const float gl = dloss != nullptr ? *dloss : 1.f;
  const float denom = mask != nullptr ? 1.f / (sums[1] + 1e-8f) : 1.f / (float)N;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
    const long long tr = t[r];
    const float* tb = tables + tr * 8;
    const float w = gl * denom * (mask != nullptr ? mask[r] : 1.f);
    for (int e = 0; e < D; ++e) {
      const float a = x0[r * D + e], nz = noise[r * D + e];
      const float xt = tb[0] * a + tb[1] * nz;
      const float eh = out[r * ldo + e], v = out[r * ldo + D + e];
      float term, dv;
      vb_element(a, xt, eh, v, tb, tr == 0, &term, &dv);
      dout[r * ldd + e] = __float2bfloat16(w * (-2.f / D) * (nz - eh));
      dout[r * ldd + D + e] = __float2bfloat16(w * dv / (D * kLn2));
    }
    for (int e = 2 * D; e < ldd; ++e) dout[r * ldd + e] = __float2bfloat16(0.f);
  }
}

// one ancestral step at spaced index `step` (all rows share it): gaussian_diffusion.py:237-314,358-392
__global__ void __launch_bounds__(128) mar_p_sample_kernel(const float* out, long long ldo, const float* x, const float* noise,
                                                           const float* tables, int step, long long N, int D,
                                                           float temperature, int clip, float* x_next, __nv_bfloat16* x16,
                                                           int kpad) {
  pdl_wait();
  const float* tb = tables + (long long)step * 8;
  const long long total = N * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kpad;
    const int e = (int)(i % kpad);
    float nx = 0.f;
    if (e < D) {
      const float xv = x[r * D + e], eps = out[r * ldo + e], v = out[r * ldo + D + e];
      const float frac = 0.5f * (v + 1.f);
      const float lv = frac * tb[7] + (1.f - frac) * tb[6];
      float px0 = tb[2] * xv - tb[3] * eps;
      if (clip) px0 = fminf(fmaxf(px0, -10.f), 10.f);  // gaussian_diffusion.py:296-298
      const float mean = tb[4] * px0 + tb[5] * xv;
      nx = mean + (step != 0 ? expf(0.5f * lv) * noise[r * D + e] * temperature : 0.f);
      x_next[r * D + e] = nx;
    }
    if (x16 != nullptr) x16[i] = __float2bfloat16(nx);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// row gather / scatter, dropout
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mar_gather_rows_kernel(const float* src, const int* idx, long long n, int C, float* dst32,
                                                              __nv_bfloat16* dst16) {
  pdl_wait();
  const long long total = n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[(long long)idx[i / C] * C + i % C];
    if (dst32 != nullptr) dst32[i] = v;
    if (dst16 != nullptr) dst16[i] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(256) mar_scatter_rows_kernel(const float* src, const int* idx, long long n, int C, float* dst) {
  pdl_wait();
  const long long total = n * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[(long long)idx[i / C] * C + i % C] = src[i];
}

__device__ __forceinline__ bool drop_keep(unsigned long long seed, long long i, uint32_t thresh) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32) >= thresh;
}

// mode 0: x16 in place; 1: out32 = resid + drop(a32); 2: out16 = bf16(drop(a32)). Four elements per thread per iteration
// (8- and 16-byte accesses); the keep decision is a function of (seed, element index) only.
__global__ void __launch_bounds__(256) dropout_kernel(int mode, __nv_bfloat16* x16, const float* a32, const float* resid,
                                                      float* out32, long long count, uint32_t thresh, float keep_scale,
                                                      unsigned long long seed, const unsigned long long* seed_dev) {
  pdl_wait();
  if (seed_dev != nullptr) seed += *seed_dev * 0xD1342543DE82EF95ull;  // per-step seed that a CUDA-graph replay can change
  const long long n4 = count / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    float k[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) k[e] = drop_keep(seed, q * 4 + e, thresh) ? keep_scale : 0.f;
    if (mode == 0) {
      uint2 v = reinterpret_cast<uint2*>(x16)[q];
      v.x = pack_bf16(bf16_lo(v.x) * k[0], bf16_hi(v.x) * k[1]);
      v.y = pack_bf16(bf16_lo(v.y) * k[2], bf16_hi(v.y) * k[3]);
      reinterpret_cast<uint2*>(x16)[q] = v;
    } else {
      const float4 a = reinterpret_cast<const float4*>(a32)[q];
      if (mode == 1) {
        const float4 r = reinterpret_cast<const float4*>(resid)[q];
        reinterpret_cast<float4*>(out32)[q] = make_float4(fmaf(a.x, k[0], r.x), fmaf(a.y, k[1], r.y), fmaf(a.z, k[2], r.z),
                                                          fmaf(a.w, k[3], r.w));
      } else {
        reinterpret_cast<uint2*>(x16)[q] = make_uint2(pack_bf16(a.x * k[0], a.y * k[1]), pack_bf16(a.z * k[2], a.w * k[3]));
      }
    }
  }
  // tail (count % 4 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 * 4; i < count; ++i) {
      const float k = drop_keep(seed, i, thresh) ? keep_scale : 0.f;
      if (mode == 0) x16[i] = __float2bfloat16(bf16_to_f(x16[i]) * k);
      else if (mode == 1) out32[i] = resid[i] + a32[i] * k;
      else x16[i] = __float2bfloat16(a32[i] * k);
    }
  }
}

static int fill_embed_dims(EmbedDims* d, int B, int T, int H, int W, int Cv, int p, int A, int pos_n) {
  HMA_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && Cv > 0 && p > 0 && A >= 0, "mar_embed: bad shape");
  HMA_REQUIRE(H % p == 0 && W % p == 0, "mar_embed: H=%d, W=%d must be multiples of the patch size %d", H, W, p);
  d->B = B; d->T = T; d->H = H; d->W = W; d->Cv = Cv; d->p = p; d->A = A; d->pos_n = pos_n;
  d->hp = H / p; d->wp = W / p; d->Sp = d->hp * d->wp; d->D = Cv * p * p; d->n = d->Sp + A;
  HMA_REQUIRE(d->D <= 64, "mar_embed: patch vector of %d elements is not supported (max 64)", d->D);
  HMA_REQUIRE(d->n <= pos_n, "mar_embed: %d tokens per frame exceed the positional table (%d)", d->n, pos_n);
  return 0;
}

}  // namespace hma

using namespace hma;
#define STREAM static_cast<cudaStream_t>(stream_)

extern "C" int hma_mar_embed_fwd(float* lat, const unsigned char* mask, const float* mask_token, const float* xp_in,
                                 const float* We, const float* act, const float* pos, int pos_n, int B, int T, int H, int W,
                                 int Cv, int p, int A, int fill_inplace, float* u, float* xp_out, float* rowmask,
                                 void* stream_) {
  EmbedDims d;
  if (int rc = fill_embed_dims(&d, B, T, H, W, Cv, p, A, pos_n)) return rc;
  HMA_REQUIRE(xp_in != nullptr || lat != nullptr, "mar_embed_fwd: no input");
  HMA_REQUIRE(A == 0 || act != nullptr, "mar_embed_fwd: action tokens requested without an action embedding");
  HMA_REQUIRE(mask == nullptr || mask_token != nullptr, "mar_embed_fwd: mask given without mask_token");
  const size_t smem = (size_t)d.D * 256 * sizeof(float);
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(mar_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 256 * 4));
    attr_done = true;
  }
  const int grid = grid_for((long long)B * T * d.n, 8, 4);
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_embed_fwd_kernel, dim3(grid), dim3(256), smem, STREAM, lat, mask, mask_token, xp_in,
                                      We, act, pos, d, fill_inplace, u, xp_out, rowmask));
  return 0;
}

extern "C" int hma_mar_embed_bwd(const float* du, const float* xp, const unsigned char* mask, const float* We, int pos_n, int B,
                                 int T, int H, int W, int Cv, int p, int A, float* dWe, float* dmask_token, float* dact,
                                 float* dpos, void* stream_) {
  EmbedDims d;
  if (int rc = fill_embed_dims(&d, B, T, H, W, Cv, p, A, pos_n)) return rc;
  HMA_REQUIRE(du != nullptr && xp != nullptr && dWe != nullptr && dpos != nullptr, "mar_embed_bwd: null argument");
  const size_t smem = ((size_t)2 * d.D * 256 + 8) * sizeof(float);
  static hma_host::PerDeviceFlag attr_flag;  // function attributes are per device (context)
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(mar_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (2 * 64 * 256 + 8) * 4));
    attr_done = true;
  }
  HMA_REQUIRE(Cv <= 8, "mar_embed_bwd: vae_embed_dim %d > 8 is not supported", Cv);
  const int grid = grid_for((long long)B * T * d.n, 64, 1);
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_embed_bwd_kernel, dim3(grid), dim3(256), smem, STREAM, du, xp, mask, We, d, dWe,
                                      dmask_token, dact, dpos));
  return 0;
}

extern "C" int hma_mar_ln_fwd(const float* x, int rows, int C, const float* gamma, const float* beta, float eps,
                              const void* mod, long long ldmod, int shift_off, int scale_off, const float* add, int add_rows,
                              float* y32, void* y16, float* stats, const void* gmod, long long ldg, int gate_off,
                              const void* h2, float* xsum, void* stream_) {
  if (rows == 0) return 0;
  HMA_REQUIRE(h2 == nullptr || (gmod != nullptr && ldg % 4 == 0 && gate_off % 4 == 0), "mar_ln_fwd: fused gate needs gmod");
  HMA_REQUIRE(C == 256 || C == 1024, "mar_ln_fwd: width %d is not supported (256 or 1024)", C);
  HMA_REQUIRE((gamma == nullptr) == (beta == nullptr), "mar_ln_fwd: gamma and beta go together");
  HMA_REQUIRE(mod == nullptr || (ldmod % 4 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0), "mar_ln_fwd: unaligned mod");
  HMA_REQUIRE(add == nullptr || add_rows > 0, "mar_ln_fwd: add_rows");
  LnArgs a{x, rows, gamma, beta, eps, static_cast<const __nv_bfloat16*>(mod), ldmod, shift_off, scale_off, add, add_rows,
           y32, static_cast<__nv_bfloat16*>(y16), stats, static_cast<const __nv_bfloat16*>(gmod), ldg, gate_off,
           static_cast<const __nv_bfloat16*>(h2), xsum};
  const int grid = grid_for(rows, 8, 8);
  if (C == 256)
    HMA_CHECK_CUDA(hma_host::launch_pdl(mar_ln_fwd_kernel<2>, dim3(grid), dim3(256), 0, STREAM, a));
  else
    HMA_CHECK_CUDA(hma_host::launch_pdl(mar_ln_fwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, a));
  return 0;
}

extern "C" int hma_mar_ln_bwd(const void* dy16, const float* dy32, const float* x, const float* stats, int rows, int C,
                              const float* gamma, const float* beta, const void* mod, long long ldmod, int shift_off,
                              int scale_off, float* dx32, int accumulate, void* dx16, float* dgamma, float* dbeta, void* dmod,
                              long long lddmod, float* dadd, int add_rows, void* stream_) {
  if (rows == 0) return 0;
  HMA_REQUIRE(C == 256 || C == 1024, "mar_ln_bwd: width %d is not supported (256 or 1024)", C);
  HMA_REQUIRE((dy16 != nullptr) != (dy32 != nullptr), "mar_ln_bwd: exactly one of dy16 / dy32");
  HMA_REQUIRE(x != nullptr && stats != nullptr, "mar_ln_bwd: x and stats are required");
  HMA_REQUIRE((gamma == nullptr) == (dgamma == nullptr) && (dgamma == nullptr) == (dbeta == nullptr) &&
                  (gamma == nullptr) == (beta == nullptr),
              "mar_ln_bwd: gamma/beta/dgamma/dbeta go together");
  HMA_REQUIRE(dmod == nullptr || mod != nullptr, "mar_ln_bwd: dmod without mod");
  HMA_REQUIRE(dadd == nullptr || add_rows > 0, "mar_ln_bwd: add_rows");
  LnBwdArgs a{static_cast<const __nv_bfloat16*>(dy16), dy32, x, stats, rows, gamma, beta,
              static_cast<const __nv_bfloat16*>(mod), ldmod, shift_off, scale_off, dx32, accumulate,
              static_cast<__nv_bfloat16*>(dx16), dgamma, dbeta, static_cast<__nv_bfloat16*>(dmod), lddmod, dadd, add_rows};
  const int grid = grid_for(rows, 8, 2);  // one resident CTA per SM at C = 1024 (registers): two full waves at most
  if (C == 256)
    HMA_CHECK_CUDA(hma_host::launch_pdl(mar_ln_bwd_kernel<2>, dim3(grid), dim3(256), 0, STREAM, a));
  else
    HMA_CHECK_CUDA(hma_host::launch_pdl(mar_ln_bwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, a));
  return 0;
}

extern "C" int hma_mar_gate_fwd(const float* x, const void* mod, long long ldmod, int gate_off, const void* h2, int rows, int C,
                                float* out, void* stream_) {
  if (rows == 0) return 0;
  HMA_REQUIRE(C % 4 == 0 && ldmod % 4 == 0 && gate_off % 4 == 0, "mar_gate_fwd: unaligned");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_gate_fwd_kernel, dim3(grid_for((long long)rows * C / 4, 256)), dim3(256), 0, STREAM, x,
                                      static_cast<const __nv_bfloat16*>(mod), ldmod, gate_off,
                                      static_cast<const __nv_bfloat16*>(h2), (long long)rows, C, out));
  return 0;
}

extern "C" int hma_mar_gate_bwd(const float* dx, const void* mod, long long ldmod, int gate_off, const void* h2, int rows, int C,
                                void* dh2, void* dmod, long long lddmod, void* stream_) {
  if (rows == 0) return 0;
  HMA_REQUIRE(C % 4 == 0 && ldmod % 4 == 0 && lddmod % 4 == 0 && gate_off % 4 == 0, "mar_gate_bwd: unaligned");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_gate_bwd_kernel, dim3(grid_for((long long)rows * C / 4, 256)), dim3(256), 0, STREAM, dx,
                                      static_cast<const __nv_bfloat16*>(mod), ldmod, gate_off,
                                      static_cast<const __nv_bfloat16*>(h2), (long long)rows, C,
                                      static_cast<__nv_bfloat16*>(dh2), static_cast<__nv_bfloat16*>(dmod), lddmod));
  return 0;
}

extern "C" int hma_mar_silu_fwd(const float* y, const float* rowvec, long long rows, int C, void* out16, void* stream_) {
  if (rows == 0) return 0;
  HMA_REQUIRE(C % 4 == 0, "mar_silu_fwd: C must be a multiple of 4");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_silu_fwd_kernel, dim3(grid_for(rows * C / 4, 256)), dim3(256), 0, STREAM, y, rowvec,
                                      rows, C, static_cast<__nv_bfloat16*>(out16)));
  return 0;
}

extern "C" int hma_mar_silu_steps(const float* c, const float* te, long long n, int steps, int C, void* out16, void* stream_) {
  if (n == 0 || steps == 0) return 0;
  HMA_REQUIRE(C % 4 == 0, "mar_silu_steps: C must be a multiple of 4");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_silu_steps_kernel, dim3(grid_for(n * steps * (C / 4), 256)), dim3(256), 0, STREAM, c, te,
                                      n, steps, C, static_cast<__nv_bfloat16*>(out16)));
  return 0;
}

extern "C" int hma_mar_silu_bwd(const float* dsy, const float* y, long long count, void* dy16, void* stream_) {
  if (count == 0) return 0;
  HMA_REQUIRE(count % 4 == 0, "mar_silu_bwd: count must be a multiple of 4");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_silu_bwd_kernel, dim3(grid_for(count / 4, 256)), dim3(256), 0, STREAM, dsy, y,
                                      count / 4, static_cast<__nv_bfloat16*>(dy16)));
  return 0;
}

extern "C" int hma_mar_q_sample(const float* x0, const float* noise, const long long* t, const float* tables, long long N, int D,
                                int kpad, void* xt16, void* stream_) {
  if (N == 0) return 0;
  HMA_REQUIRE(kpad >= D, "mar_q_sample: kpad < D");
  HMA_REQUIRE(noise == nullptr || (t != nullptr && tables != nullptr), "mar_q_sample: noise needs t and tables");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_q_sample_kernel, dim3(grid_for(N * kpad, 256)), dim3(256), 0, STREAM, x0, noise, t,
                                      tables, N, D, kpad, static_cast<__nv_bfloat16*>(xt16)));
  return 0;
}

extern "C" int hma_mar_timestep_embed(const long long* t, long long N, void* out16, void* stream_) {
  if (N == 0) return 0;
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_timestep_embed_kernel, dim3(grid_for(N * 256, 256)), dim3(256), 0, STREAM, t, N,
                                      static_cast<__nv_bfloat16*>(out16)));
  return 0;
}

extern "C" int hma_mar_diff_loss_fwd(const float* out, long long ldo, const float* x0, const float* noise, const long long* t,
                                     const float* mask, const float* tables, long long N, int D, float* rows_loss, float* sums,
                                     float* loss, void* stream_) {
  HMA_REQUIRE(N > 0 && D > 0 && ldo >= 2 * D, "mar_diff_loss_fwd: bad shape");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_diff_loss_fwd_kernel, dim3(grid_for(N, 128)), dim3(128), 0, STREAM, out, ldo, x0, noise,
                                      t, mask, tables, N, D, rows_loss, sums));
  if (loss != nullptr)
    HMA_CHECK_CUDA(hma_host::launch_pdl(mar_diff_loss_finish_kernel, dim3(1), dim3(1), 0, STREAM, (const float*)sums, N,
                                        mask != nullptr ? 1 : 0, loss));
  return 0;
}

extern "C" int hma_mar_diff_loss_bwd(const float* out, long long ldo, const float* x0, const float* noise, const long long* t,
                                     const float* mask, const float* tables, long long N, int D, const float* sums,
                                     const float* dloss, void* dout16, long long ldd, void* stream_) {
  HMA_REQUIRE(N > 0 && D > 0 && ldo >= 2 * D && ldd >= 2 * D, "mar_diff_loss_bwd: bad shape");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_diff_loss_bwd_kernel, dim3(grid_for(N, 128)), dim3(128), 0, STREAM, out, ldo, x0, noise,
                                      t, mask, tables, N, D, sums, dloss, static_cast<__nv_bfloat16*>(dout16), ldd));
  return 0;
}

extern "C" int hma_mar_p_sample(const float* out, long long ldo, const float* x, const float* noise, const float* tables,
                                int step, long long N, int D, float temperature, int clip, float* x_next, void* x16, int kpad,
                                void* stream_) {
  if (N == 0) return 0;
  HMA_REQUIRE(D > 0 && ldo >= 2 * D && kpad >= D && step >= 0, "mar_p_sample: bad shape");
  HMA_REQUIRE(step == 0 || noise != nullptr, "mar_p_sample: noise is required for step > 0");
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_p_sample_kernel, dim3(grid_for(N * kpad, 128)), dim3(128), 0, STREAM, out, ldo, x,
                                      noise, tables, step, N, D, temperature, clip, x_next,
                                      static_cast<__nv_bfloat16*>(x16), kpad));
  return 0;
}

extern "C" int hma_mar_gather_rows(const float* src, const int* idx, long long n, int C, float* dst32, void* dst16,
                                   void* stream_) {
  if (n == 0) return 0;
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_gather_rows_kernel, dim3(grid_for(n * C, 256)), dim3(256), 0, STREAM, src, idx, n, C,
                                      dst32, static_cast<__nv_bfloat16*>(dst16)));
  return 0;
}

extern "C" int hma_mar_scatter_rows(const float* src, const int* idx, long long n, int C, float* dst, void* stream_) {
  if (n == 0) return 0;
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_scatter_rows_kernel, dim3(grid_for(n * C, 256)), dim3(256), 0, STREAM, src, idx, n, C,
                                      dst));
  return 0;
}

static int launch_dropout(int mode, void* x16, const float* a32, const float* resid, float* out32, long long count, float p,
                          unsigned long long seed, const unsigned long long* seed_dev, cudaStream_t stream) {
  if (count == 0) return 0;
  HMA_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f out of range", p);
  const uint32_t thresh = (uint32_t)fmin(4294967295.0, (double)p * 4294967296.0);
  HMA_REQUIRE(((reinterpret_cast<uintptr_t>(x16) | reinterpret_cast<uintptr_t>(a32) | reinterpret_cast<uintptr_t>(resid) |
                reinterpret_cast<uintptr_t>(out32)) & 15) == 0, "dropout: buffers must be 16-byte aligned");
  HMA_CHECK_CUDA(hma_host::launch_pdl(dropout_kernel, dim3(grid_for((count + 3) / 4, 256)), dim3(256), 0, stream, mode,
                                      static_cast<__nv_bfloat16*>(x16), a32, resid, out32, count, thresh, 1.f / (1.f - p),
                                      seed, seed_dev));
  return 0;
}

extern "C" int hma_dropout_bf16(void* x, long long count, float p, unsigned long long seed,
                                const unsigned long long* seed_dev, void* stream_) {
  return launch_dropout(0, x, nullptr, nullptr, nullptr, count, p, seed, seed_dev, STREAM);
}
extern "C" int hma_dropout_add_f32(const float* a, const float* resid, float* out, long long count, float p,
                                   unsigned long long seed, const unsigned long long* seed_dev, void* stream_) {
  return launch_dropout(1, nullptr, a, resid, out, count, p, seed, seed_dev, STREAM);
}
extern "C" int hma_dropout_cast_bf16(const float* a, void* out16, long long count, float p, unsigned long long seed,
                                     const unsigned long long* seed_dev, void* stream_) {
  return launch_dropout(2, out16, a, nullptr, nullptr, count, p, seed, seed_dev, STREAM);
}
