// Frame-incremental causal temporal attention for MaskGIT decode, and the K/V cache it reads.
//
// The reference recomputes the whole T-frame window for every MaskGIT step of every generated frame
// (st_mask_git.py:384,394: compute_logits(prompt_THW) inside the step loop). Temporal attention is
// causal (st_transformer.py:111) and spatial attention is per frame, so frames < out_t never change
// while frame out_t is being decoded: their per-layer temporal keys/values are kept in a cache
//     kv[frame][b * n + s][2 * C]      (K in columns [0, C), V in [C, 2C); bf16, frame-major)
// and a decode step only runs the ONE frame being generated through the network. This kernel is the
// temporal attention of that frame: the query of token (b, s) attends to the cached K/V of the same
// slot in frames [0, n_prev) plus its own K/V (causal "<=").
// A pass may also carry `frames` consecutive window frames per sample ((b, f, s) row order; their K/V are appended to the
// cache BEFORE this kernel runs): frame f then attends to cache frames [0, n_prev + f) plus itself — how a finished frame
// (f = 0) is committed in the same pass that runs the first MaskGIT step of the next one (f = 1).
//
// HBM-bound by the cache read (n_prev * 1 KB per token), so it runs on the CUDA cores: one warp per
// token, lane l owns channels [8l, 8l+8) of all 8 heads' 256 channels (head = l / 4), every K/V row
// is one fully coalesced 512-byte warp load, scores are reduced over the 4 lanes of a head and the
// softmax is the usual single-pass online form in fp32.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kTC = 256;  // channels (8 heads x 32)

struct TemporalCachedParams {
  const __nv_bfloat16* qkv;  // [rows, ld_qkv]: this frame's q | k | v
  long long ld_qkv;
  int q_col, k_col, v_col;
  const __nv_bfloat16* kv;   // cache base
  long long frame_stride;    // elements between consecutive frames of the cache
  int rows;                  // B * frames * n tokens of the pass
  int n_prev;                // cached frames the first frame of the pass attends to
  int frames, n;             // frames per sample in this pass; tokens per frame
  float scale_log2;
  __nv_bfloat16* out;        // [rows, ldo]
  long long ldo;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

__device__ __forceinline__ float head_dot(const float (&q)[8], const uint4& ku) {
  float k[8];
  unpack8(ku, k);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s = fmaf(q[j], k[j], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}

__global__ void __launch_bounds__(256) attn_temporal_cached_kernel(const TemporalCachedParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  pdl_wait();
  pdl_launch_dependents();
  if (row >= p.rows) return;
  const __nv_bfloat16* qrow = p.qkv + row * p.ld_qkv;
  float q[8];
  unpack8(*reinterpret_cast<const uint4*>(qrow + p.q_col + lane * 8), q);
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] *= p.scale_log2;
  const uint4 k_own = *reinterpret_cast<const uint4*>(qrow + p.k_col + lane * 8);
  const uint4 v_own = *reinterpret_cast<const uint4*>(qrow + p.v_col + lane * 8);

  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto fold = [&](const uint4& ku, const uint4& vu) {
    const float s = head_dot(q, ku);
    const float mn = fmaxf(m, s);
    const float corr = fast_ex2(m - mn);  // first key: ex2(-inf) = 0
    const float pe = fast_ex2(s - mn);
    float v[8];
    unpack8(vu, v);
    l = fmaf(l, corr, pe);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(acc[j], corr, pe * v[j]);
    m = mn;
  };

  long long crow = row;  // row of this token's slot in a cache frame: b * n + s
  int n_prev = p.n_prev;
  if (p.frames > 1) {
    const long long bf = row / p.n;
    crow = (bf / p.frames) * p.n + (row - bf * p.n);
    n_prev += (int)(bf % p.frames);
  }
  const __nv_bfloat16* base = p.kv + crow * (2 * kTC) + lane * 8;
  int f = 0;
  for (; f + 4 <= n_prev; f += 4) {  // four frames of loads in flight per lane
    uint4 ku[4], vu[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat16* r = base + (long long)(f + u) * p.frame_stride;
      ku[u] = __ldg(reinterpret_cast<const uint4*>(r));
      vu[u] = __ldg(reinterpret_cast<const uint4*>(r + kTC));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) fold(ku[u], vu[u]);
  }
  for (; f < n_prev; ++f) {
    const __nv_bfloat16* r = base + (long long)f * p.frame_stride;
    fold(__ldg(reinterpret_cast<const uint4*>(r)), __ldg(reinterpret_cast<const uint4*>(r + kTC)));
  }
  fold(k_own, v_own);

  const float inv = 1.0f / l;
  *reinterpret_cast<uint4*>(p.out + row * p.ldo + lane * 8) =
      make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                 pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
}

// K/V columns of `frames` frames of a (b, t, s)-ordered qkv matrix -> cache frames [t0, t0 + frames).
struct KvAppendParams {
  const __nv_bfloat16* qkv;
  long long ld_qkv;
  int k_col, v_col;
  int B, frames, n;
  __nv_bfloat16* kv;
  long long frame_stride;
  int t0;
};

__global__ void __launch_bounds__(256) kv_append_kernel(const KvAppendParams p) {
  const long long total = (long long)p.B * p.frames * p.n * 64;  // 16-byte pieces: 32 of K + 32 of V per token
  pdl_wait();
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int piece = (int)(i & 63);
    const long long tok = i >> 6;  // (b, t, s) order
    const int s = (int)(tok % p.n);
    const long long bt = tok / p.n;
    const int t = (int)(bt % p.frames);
    const int b = (int)(bt / p.frames);
    const int col = piece < 32 ? p.k_col + piece * 8 : p.v_col + (piece - 32) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(p.qkv + tok * p.ld_qkv + col);
    *reinterpret_cast<uint4*>(p.kv + (long long)(p.t0 + t) * p.frame_stride + ((long long)b * p.n + s) * (2 * kTC) + piece * 8) = v;
  }
}

}  // namespace hma

extern "C" int hma_attn_temporal_cached(const void* qkv, long long ld_qkv, int q_col, int k_col, int v_col, const void* kv,
                                        long long frame_stride, int rows, int n_prev, int heads, float scale, void* out,
                                        long long ldo, int frames, int n, void* stream_) {
  using namespace hma;
  if (rows == 0) return 0;
  HMA_REQUIRE(heads == 8, "attn_temporal_cached: built for 8 heads of 32 channels (got %d heads)", heads);
  HMA_REQUIRE(n_prev >= 0 && (n_prev == 0 || kv != nullptr), "attn_temporal_cached: bad cache arguments");
  HMA_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 && frame_stride % 8 == 0,
              "attn_temporal_cached: 16-byte alignment required");
  TemporalCachedParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv); p.ld_qkv = ld_qkv;
  p.q_col = q_col; p.k_col = k_col; p.v_col = v_col;
  p.kv = static_cast<const __nv_bfloat16*>(kv); p.frame_stride = frame_stride;
  p.rows = rows; p.n_prev = n_prev;
  HMA_REQUIRE(frames >= 1 && (frames == 1 || (n > 0 && rows % (frames * n) == 0)),
              "attn_temporal_cached: rows=%d is not a whole number of %d-frame samples of %d tokens", rows, frames, n);
  p.frames = frames; p.n = n > 0 ? n : 1;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  HMA_CHECK_CUDA(hma_host::launch_pdl(attn_temporal_cached_kernel, dim3((rows + 7) / 8), dim3(256), 0,
                                      static_cast<cudaStream_t>(stream_), p));
  return 0;
}

extern "C" int hma_kv_cache_append(const void* qkv, long long ld_qkv, int k_col, int v_col, int B, int frames, int n,
                                   void* kv, long long frame_stride, int t0, void* stream_) {
  using namespace hma;
  if (B == 0 || frames == 0 || n == 0) return 0;
  HMA_REQUIRE(ld_qkv % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 && frame_stride % 8 == 0,
              "kv_cache_append: 16-byte alignment required");
  HMA_REQUIRE(frame_stride >= (long long)B * n * 2 * kTC, "kv_cache_append: frame stride smaller than one frame");
  KvAppendParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv); p.ld_qkv = ld_qkv; p.k_col = k_col; p.v_col = v_col;
  p.B = B; p.frames = frames; p.n = n;
  p.kv = static_cast<__nv_bfloat16*>(kv); p.frame_stride = frame_stride; p.t0 = t0;
  const long long total = (long long)B * frames * n * 64;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)hma_host::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  HMA_CHECK_CUDA(hma_host::launch_pdl(kv_append_kernel, dim3((int)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream_), p));
  return 0;
}
