// Token -> pixel decode (reference: hma/visualize.py:136-151, external/magvit2
// lookup_free_quantize.py:181-194 and improved_model.py:12-51,124-234): the element-wise stages around the 3x3
// convolutions, which are tcgen05 contractions over zero-bordered NHWC images (hma_conv3x3_nhwc, gemm_nt.cu).
//
// Layout: an activation of an image batch is [images * (H+2) * (W+2), C] — NHWC with a one-pixel zero border per image, so
// that the nine taps of a 3x3 convolution are nine row offsets of the same matrix. fp32 for what the convolutions
// accumulate into (the residual stream of the decoder), bf16 for their operands. All HBM-bound:
//   lfq_entry        token ids -> +-1 code bits, bf16, border and channel padding zero                (visualize.py:149-150)
//   gn_stats         per (image, group) sum and sum of squares of the interior pixels, fixed order    (nn.GroupNorm(32, C, eps=1e-6))
//   gn_swish         (x - mean) * rstd * gamma + beta -> x * sigmoid(x) -> bf16, border zero          (improved_model.py:38-44,178-179)
//   cast_bordered    fp32 -> bf16 with the border forced to zero (operand of the 1x1 nin_shortcut)
//   depth_to_space   [.., 4C'] at (h, w) -> [.., C'] at (2h+i, 2w+j), channel (i*2+j)*C' + c (DCR)     (improved_model.py:185-217)
//   to_uint8         clamp(-1, 1) -> (x + 1) * 127.5 -> clamp(0, 255) -> truncate, NCHW uint8          (visualize.py:112-121)
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

__device__ __forceinline__ bool interior(int pix, int Hp, int Wp, int& y, int& x) {
  y = pix / Wp;
  x = pix - y * Wp;
  return y >= 1 && y < Hp - 1 && x >= 1 && x < Wp - 1;
}

// one thread per (padded pixel, 8 channels)
__global__ void __launch_bounds__(256) lfq_entry_kernel(const long long* tokens, int images, int H, int W, int bits, int ldc,
                                                        __nv_bfloat16* out) {
  pdl_wait();
  pdl_launch_dependents();
  const int Hp = H + 2, Wp = W + 2, vec = ldc / 8;
  const long long total = (long long)images * Hp * Wp * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long prow = i / vec;
    const int img = (int)(prow / (Hp * Wp));
    int y, x;
    const bool in = interior((int)(prow - (long long)img * Hp * Wp), Hp, Wp, y, x);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (in) {
      const long long id = tokens[((long long)img * H + (y - 1)) * W + (x - 1)];
      // get_codebook_entry writes big-endian bits (channel c = bit bits-1-c) and visualize.py flips the channel axis:
      // channel c of the decoder input is bit c of the id
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int c = v * 8 + j;
        const float b0 = c < bits ? (((id >> c) & 1) ? 1.f : -1.f) : 0.f;
        const float b1 = c + 1 < bits ? (((id >> (c + 1)) & 1) ? 1.f : -1.f) : 0.f;
        w[j >> 1] = pack_bf16(b0, b1);
      }
    }
    *reinterpret_cast<uint4*>(out + prow * ldc + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// GroupNorm statistics in a FIXED summation order (bit-reproducible frames): grid = (chunks, images); a thread owns 4
// consecutive channels (inside one group: C / 32 >= 4) over a strided set of pixels; the CTA's partial (sum, sum of squares)
// per group is reduced in thread order and written to scratch[image, chunk, 32, 2]; gn_finalize adds the chunks in order.
constexpr int kGnMaxChunks = 64;
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* x, int H, int W, int C, float* scratch) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float part[256][2];
  const int Wp = W + 2, Hp = H + 2, vec = C / 4, cpg = C / 32;
  const int img = blockIdx.y;
  const float* base = x + (size_t)img * Hp * Wp * C;
  const int v = threadIdx.x % vec;            // blockDim.x % vec == 0 (C in {128, 256, 512} -> vec in {32, 64, 128})
  const int lanes = blockDim.x / vec;         // pixels processed concurrently by the CTA
  float s = 0.f, ss = 0.f;
  for (int pix = blockIdx.x * lanes + threadIdx.x / vec; pix < H * W; pix += gridDim.x * lanes) {
    const int y = pix / W + 1, xx = pix % W + 1;
    const float4 t = *reinterpret_cast<const float4*>(base + (size_t)(y * Wp + xx) * C + v * 4);
    s += t.x + t.y + t.z + t.w;
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  part[threadIdx.x][0] = s;
  part[threadIdx.x][1] = ss;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int grp = threadIdx.x >> 1, which = threadIdx.x & 1;
    const int v_per_grp = cpg / 4;
    float acc = 0.f;
    for (int lane = 0; lane < lanes; ++lane)
      for (int vv = grp * v_per_grp; vv < (grp + 1) * v_per_grp; ++vv) acc += part[lane * vec + vv][which];
    scratch[((size_t)img * gridDim.x + blockIdx.x) * 64 + threadIdx.x] = acc;
  }
}
__global__ void __launch_bounds__(64) gn_finalize_kernel(const float* scratch, int chunks, float* sums) {
  pdl_wait();
  pdl_launch_dependents();
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += scratch[((size_t)blockIdx.x * chunks + c) * 64 + threadIdx.x];
  sums[(size_t)blockIdx.x * 64 + threadIdx.x] = acc;
}

// mode 0: GroupNorm + swish; mode 1: plain cast. One thread per (padded pixel, 8 channels). Border pixels -> zero.
__global__ void __launch_bounds__(256) gn_swish_kernel(const float* x, const float* sums, const float* gamma, const float* beta,
                                                       int images, int H, int W, int C, float eps, int mode, __nv_bfloat16* out) {
  pdl_wait();
  pdl_launch_dependents();
  const int Hp = H + 2, Wp = W + 2, vec = C / 8, cpg = C / 32;
  const float inv_n = 1.0f / ((float)H * (float)W * (float)cpg);
  const long long total = (long long)images * Hp * Wp * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long prow = i / vec;
    const int img = (int)(prow / (Hp * Wp));
    int y, xx;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (interior((int)(prow - (long long)img * Hp * Wp), Hp, Wp, y, xx)) {
      const float4 a = *reinterpret_cast<const float4*>(x + prow * C + v * 8);
      const float4 b = *reinterpret_cast<const float4*>(x + prow * C + v * 8 + 4);
      float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = v * 8 + j, grp = c / cpg;
          const float mean = sums[(size_t)img * 64 + 2 * grp] * inv_n;
          const float var = fmaxf(sums[(size_t)img * 64 + 2 * grp + 1] * inv_n - mean * mean, 0.f);
          const float h = (t[j] - mean) * rsqrtf(var + eps) * gamma[c] + beta[c];
          t[j] = h / (1.0f + __expf(-h));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) w[j >> 1] = pack_bf16(t[j], t[j + 1]);
    }
    *reinterpret_cast<uint4*>(out + prow * C + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// in: fp32 [images, (H+2)(W+2), 4*Co] (interior valid); out: fp32 [images, (2H+2)(2W+2), Co], border zero.
// One thread per (padded output pixel, 4 channels).
__global__ void __launch_bounds__(256) depth_to_space_kernel(const float* in, int images, int H, int W, int Co, float* out) {
  pdl_wait();
  pdl_launch_dependents();
  const int Hp = H + 2, Wp = W + 2, Ho = 2 * H + 2, Wo = 2 * W + 2, vec = Co / 4;
  const long long total = (long long)images * Ho * Wo * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long prow = i / vec;
    const int img = (int)(prow / (Ho * Wo));
    int y, x;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (interior((int)(prow - (long long)img * Ho * Wo), Ho, Wo, y, x)) {
      const int oy = y - 1, ox = x - 1;
      const int h = oy >> 1, ii = oy & 1, w = ox >> 1, jj = ox & 1;
      const size_t src = ((size_t)img * Hp * Wp + (size_t)(h + 1) * Wp + (w + 1)) * (size_t)(4 * Co) + (size_t)(ii * 2 + jj) * Co + v * 4;
      t = *reinterpret_cast<const float4*>(in + src);
    }
    *reinterpret_cast<float4*>(out + prow * Co + v * 4) = t;
  }
}

// x: fp32 [images, (H+2)(W+2), ldc] (channels 0..ch-1 used) -> uint8 [images, ch, H, W]
__global__ void __launch_bounds__(256) to_uint8_kernel(const float* x, int images, int H, int W, int ldc, int ch, unsigned char* out) {
  pdl_wait();
  pdl_launch_dependents();
  const int Wp = W + 2, Hp = H + 2;
  const long long total = (long long)images * ch * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % ch);
    const int img = (int)(i / ((long long)W * H * ch));
    float v = x[((size_t)img * Hp * Wp + (size_t)(y + 1) * Wp + (xx + 1)) * ldc + c];
    v = fminf(fmaxf(v, -1.f), 1.f);
    v = fminf(fmaxf((v + 1.f) * 127.5f, 0.f), 255.f);
    out[i] = (unsigned char)v;  // truncation, as torch's .to(uint8)
  }
}

static inline unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)hma_host::sm_count() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace hma

extern "C" int hma_lfq_entry(const long long* tokens, int images, int H, int W, int bits, int ldc, void* out, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(bits >= 1 && bits <= 62 && ldc % 8 == 0 && ldc >= bits, "lfq_entry: bad bits=%d / channel stride=%d", bits, ldc);
  if (images == 0) return 0;
  const long long total = (long long)images * (H + 2) * (W + 2) * (ldc / 8);
  HMA_CHECK_CUDA(hma_host::launch_pdl(lfq_entry_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream_), tokens,
                                      images, H, W, bits, ldc, static_cast<__nv_bfloat16*>(out)));
  return 0;
}

extern "C" int hma_gn_stats(const float* x, int images, int H, int W, int C, float* scratch, float* sums, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(C == 128 || C == 256 || C == 512, "gn_stats: C=%d must be 128, 256 or 512 (32 groups)", C);
  HMA_REQUIRE(scratch != nullptr && sums != nullptr, "gn_stats: scratch [images, 64, 64] and sums [images, 32, 2] are required");
  if (images == 0) return 0;
  int chunks = (H * W + 255) / 256;
  if (chunks > kGnMaxChunks) chunks = kGnMaxChunks;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HMA_CHECK_CUDA(hma_host::launch_pdl(gn_stats_kernel, dim3(chunks, images), dim3(256), 0, stream, x, H, W, C, scratch));
  HMA_CHECK_CUDA(hma_host::launch_pdl(gn_finalize_kernel, dim3(images), dim3(64), 0, stream, static_cast<const float*>(scratch), chunks, sums));
  return 0;
}

extern "C" int hma_gn_swish(const float* x, const float* sums, const float* gamma, const float* beta, int images, int H, int W, int C,
                            float eps, int mode, void* out, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(C % 32 == 0, "gn_swish: C=%d must be a multiple of 32", C);
  HMA_REQUIRE(mode == 1 || (sums != nullptr && gamma != nullptr && beta != nullptr), "gn_swish: GroupNorm mode needs sums, gamma, beta");
  if (images == 0) return 0;
  const long long total = (long long)images * (H + 2) * (W + 2) * (C / 8);
  HMA_CHECK_CUDA(hma_host::launch_pdl(gn_swish_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, sums,
                                      gamma, beta, images, H, W, C, eps, mode, static_cast<__nv_bfloat16*>(out)));
  return 0;
}

extern "C" int hma_depth_to_space(const float* in, int images, int H, int W, int Co, float* out, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(Co % 4 == 0, "depth_to_space: Co=%d must be a multiple of 4", Co);
  if (images == 0) return 0;
  const long long total = (long long)images * (2 * H + 2) * (2 * W + 2) * (Co / 4);
  HMA_CHECK_CUDA(hma_host::launch_pdl(depth_to_space_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream_), in,
                                      images, H, W, Co, out));
  return 0;
}

extern "C" int hma_to_uint8(const float* x, int images, int H, int W, int ldc, int ch, void* out, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(ch >= 1 && ch <= ldc, "to_uint8: bad channel count %d (stride %d)", ch, ldc);
  if (images == 0) return 0;
  const long long total = (long long)images * ch * H * W;
  HMA_CHECK_CUDA(hma_host::launch_pdl(to_uint8_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream_), x, images,
                                      H, W, ldc, ch, static_cast<unsigned char*>(out)));
  return 0;
}
