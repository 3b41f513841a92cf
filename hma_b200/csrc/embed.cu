// Token embedding + action-token concat + positional embedding in one pass, and its backward.
// Reference: factorization_utils.py:31-54,57-68 (sum of factored embeddings, mask id -> the
// mask_token_embed row), st_mask_git.py:651-661 (64 replicated action tokens appended to each
// frame) and :670-672 (+ pos_embed_TSC[:, :T, :n]). The reference does this with boolean-mask
// indexing (a nonzero() sync), two gathers, a stack+sum, a repeat, a concat and an add; here one
// warp produces one 256-float row of the fp32 residual stream directly.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kEC = 256;

struct EmbedParams {
  const long long* ids;  // [B*T*S]
  const float* E0;       // [vs, C]
  const float* E1;       // [vs, C] (null when num_factored_vocabs == 1)
  const float* mask_embed;  // [C]
  const float* act;      // [B*T, C] action embedding or null
  const float* pos;      // pos_embed_TSC, row (t, s) at pos + (t*pos_n + s)*C
  int pos_n;
  int B, T, S, A;        // A action tokens per frame (0 if none)
  int vs;
  long long mask_id;
  float* x;              // [B*T*(S+A), C]
};

__global__ void __launch_bounds__(256) embed_fwd_kernel(const EmbedParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = p.S + p.A;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)p.B * p.T * n) return;
  const int s = (int)(row % n);
  const long long bt = row / n;
  const int t = (int)(bt % p.T);
  const float* src0;
  const float* src1 = nullptr;
  if (s < p.S) {
    const long long id = p.ids[bt * p.S + s];
    if (id == p.mask_id) {
      src0 = p.mask_embed;
    } else {
      src0 = p.E0 + (size_t)(id % p.vs) * kEC;
      if (p.E1 != nullptr) src1 = p.E1 + (size_t)((id / p.vs) % p.vs) * kEC;
    }
  } else {
    src0 = p.act + (size_t)bt * kEC;
  }
  const float* pr = p.pos + ((size_t)t * p.pos_n + s) * kEC;
  float* dst = p.x + (size_t)row * kEC;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    float4 v = *reinterpret_cast<const float4*>(src0 + c);
    if (src1 != nullptr) {
      const float4 w = *reinterpret_cast<const float4*>(src1 + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c));
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    *reinterpret_cast<float4*>(dst + c) = v;
  }
}

struct EmbedBwdParams {
  EmbedParams f;
  const float* dx;   // [B*T*n, C]
  float* dE0;
  float* dE1;
  float* dmask;
  float* dact;       // [B*T, C] (accumulated) or null
  float* dpos;       // same layout as pos (accumulated)
};

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One CTA per (t, 8 consecutive s); warp w owns slot s0+w and loops over the batch, so the
// positional gradient needs no atomics and the mask-embedding gradient is reduced per CTA.
__global__ void __launch_bounds__(256) embed_bwd_kernel(const EmbedBwdParams p) {
  __shared__ float red[8][kEC];
  const EmbedParams& f = p.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = f.S + f.A;
  const int chunks = (n + 7) / 8;
  const int t = blockIdx.x / chunks;
  const int s = (blockIdx.x % chunks) * 8 + warp;
  float4 pos_acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  float4 mask_acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (s < n) {
    for (int b = 0; b < f.B; ++b) {
      const long long bt = (long long)b * f.T + t;
      const float* g = p.dx + ((size_t)bt * n + s) * kEC;
      float4 v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        v[h] = *reinterpret_cast<const float4*>(g + h * 128 + lane * 4);
        pos_acc[h].x += v[h].x; pos_acc[h].y += v[h].y; pos_acc[h].z += v[h].z; pos_acc[h].w += v[h].w;
      }
      if (s < f.S) {
        const long long id = f.ids[bt * f.S + s];
        if (id == f.mask_id) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mask_acc[h].x += v[h].x; mask_acc[h].y += v[h].y; mask_acc[h].z += v[h].z; mask_acc[h].w += v[h].w;
          }
        } else {
          float* d0 = p.dE0 + (size_t)(id % f.vs) * kEC;
#pragma unroll
          for (int h = 0; h < 2; ++h) red_add_v4(d0 + h * 128 + lane * 4, v[h]);
          if (p.dE1 != nullptr) {
            float* d1 = p.dE1 + (size_t)((id / f.vs) % f.vs) * kEC;
#pragma unroll
            for (int h = 0; h < 2; ++h) red_add_v4(d1 + h * 128 + lane * 4, v[h]);
          }
        }
      } else if (p.dact != nullptr) {
        float* d = p.dact + (size_t)bt * kEC;
#pragma unroll
        for (int h = 0; h < 2; ++h) red_add_v4(d + h * 128 + lane * 4, v[h]);
      }
    }
    float* dp = p.dpos + ((size_t)t * f.pos_n + s) * kEC;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 cur = *reinterpret_cast<float4*>(dp + h * 128 + lane * 4);
      cur.x += pos_acc[h].x; cur.y += pos_acc[h].y; cur.z += pos_acc[h].z; cur.w += pos_acc[h].w;
      *reinterpret_cast<float4*>(dp + h * 128 + lane * 4) = cur;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&red[warp][h * 128 + lane * 4]) = mask_acc[h];
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
  if (sum != 0.f) atomicAdd(p.dmask + threadIdx.x, sum);
}

}  // namespace hma

extern "C" int hma_embed_fwd(const long long* ids, const float* E0, const float* E1, const float* mask_embed,
                             const float* act, const float* pos, int pos_n, int B, int T, int S, int A, int vs,
                             long long mask_id, float* x, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(A == 0 || act != nullptr, "embed_fwd: action tokens requested without an action embedding");
  HMA_REQUIRE(S + A <= pos_n, "embed_fwd: %d tokens per frame exceed the positional table (%d)", S + A, pos_n);
  EmbedParams p{ids, E0, E1, mask_embed, act, pos, pos_n, B, T, S, A, vs, mask_id, x};
  const long long rows = (long long)B * T * (S + A);
  if (rows == 0) return 0;
  embed_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_embed_bwd(const long long* ids, const float* dx, int pos_n, int B, int T, int S, int A, int vs,
                             long long mask_id, float* dE0, float* dE1, float* dmask, float* dact, float* dpos,
                             void* stream_) {
  using namespace hma;
  EmbedBwdParams p{};
  p.f.ids = ids; p.f.pos_n = pos_n; p.f.B = B; p.f.T = T; p.f.S = S; p.f.A = A; p.f.vs = vs; p.f.mask_id = mask_id;
  p.dx = dx; p.dE0 = dE0; p.dE1 = dE1; p.dmask = dmask; p.dact = dact; p.dpos = dpos;
  const int n = S + A;
  if (B * T * n == 0) return 0;
  embed_bwd_kernel<<<T * ((n + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
