// Optimizer step over the flat parameter arena (reference:
// train_multi.py:593-598 clip_grad_norm_ + AdamW). Parameters, gradients and both moments are
// contiguous fp32 ranges with identical layout, so the step is two streaming kernels:
//   sumsq      : sum of squares of a gradient range added to a device scalar, in a fixed order (replicas stay bit-identical)
//   adamw_step : p, m, v updated in place; the global-norm clip coefficient is computed on the
//                device from that scalar, so no host synchronisation is needed
// Only the ranges that received gradients this step (shared trunk + the active domain) are touched.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

// Deterministic: every block writes its partial sum to a fixed slot and a second, single-block launch adds the slots in
// index order. (An atomicAdd of the partials summed in arrival order, so the clip coefficient — and with it every
// parameter — differed in the last bit between data-parallel replicas holding bit-identical gradients, and the
// replicas drifted apart; torch's clip_grad_norm_ in the reference trainer is deterministic.)
constexpr int kSumsqMaxBlocks = 148 * 8;
__device__ float g_sumsq_partial[kSumsqMaxBlocks];

__global__ void __launch_bounds__(256) sumsq_kernel(const float* g, long long n4) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w];
    g_sumsq_partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) sumsq_finalize_kernel(int blocks, float* out) {
  __shared__ float part[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += g_sumsq_partial[i];  // fixed assignment, fixed order
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out += part[0];  // ranges are accumulated by successive calls on one stream
}

struct AdamParams {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n4, n4_decay;
  float lr, beta1, beta2, eps, wd, bc1, bc2;  // bc = 1 - beta^t
  float grad_scale;                            // e.g. 1/world_size
  const float* sumsq;                          // device scalar (may be null: no clipping)
  float max_norm;
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
  float clip = 1.f;
  if (a.sumsq != nullptr) {
    const float norm = sqrtf(*a.sumsq) * a.grad_scale;
    clip = fminf(1.f, a.max_norm / (norm + 1e-6f));  // torch.nn.utils.clip_grad_norm_
  }
  const float gs = a.grad_scale * clip;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n4) return;
  float4 p = reinterpret_cast<float4*>(a.p)[i];
  const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
  float4 m = reinterpret_cast<float4*>(a.m)[i];
  float4 v = reinterpret_cast<float4*>(a.v)[i];
  const float decay = i < a.n4_decay ? 1.f - a.lr * a.wd : 1.f;
  float* pp = reinterpret_cast<float*>(&p);
  const float* gg = reinterpret_cast<const float*>(&g4);
  float* mm = reinterpret_cast<float*>(&m);
  float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float g = gg[j] * gs;
    pp[j] *= decay;  // decoupled weight decay (1 for the no-decay group)
    mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * g;
    vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(vv[j] / a.bc2) + a.eps;
    pp[j] -= a.lr * (mm[j] / a.bc1) / denom;
  }
  reinterpret_cast<float4*>(a.p)[i] = p;
  reinterpret_cast<float4*>(a.m)[i] = m;
  reinterpret_cast<float4*>(a.v)[i] = v;
}

}  // namespace hma

extern "C" int hma_sumsq(const float* g, long long n, float* out, void* stream_) {
  using namespace hma;
  if (n == 0) return 0;
  HMA_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: range must be 16-byte aligned");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > kSumsqMaxBlocks) blocks = kSumsqMaxBlocks;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(g, n4);
  sumsq_finalize_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>((int)blocks, out);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_adamw_step(float* p, const float* g, float* m, float* v, long long n, long long n_decay, float lr,
                              float beta1, float beta2, float eps, float wd, int step, float grad_scale,
                              const float* sumsq, float max_norm, void* stream_) {
  using namespace hma;
  if (n == 0) return 0;
  HMA_REQUIRE(n % 4 == 0, "adamw: range length must be a multiple of 4");
  HMA_REQUIRE(step >= 1, "adamw: step counts from 1");
  HMA_REQUIRE(n_decay >= 0 && n_decay <= n && n_decay % 4 == 0, "adamw: n_decay must be a multiple of 4 in [0, n]");
  AdamParams a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.n4 = n / 4; a.n4_decay = n_decay / 4;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = wd;
  a.bc1 = 1.f - powf(beta1, (float)step);
  a.bc2 = 1.f - powf(beta2, (float)step);
  a.grad_scale = grad_scale; a.sumsq = sumsq; a.max_norm = max_norm;
  adamw_kernel<<<(unsigned)((a.n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
