// Factorised cross-entropy with label smoothing, accuracy and masked mean, and its backward.
// Reference: st_mask_git.py:603-630 (compute_video_loss_and_acc), factorization_utils.py:85-96
// (labels -> per-vocabulary digits), forward :714-716 (relevant mask = input token is the mask id,
// frames 1..T-1 only).
//   per token: sum_k CE_smooth(logits[k*vs:(k+1)*vs], digit_k(label))  ;  acc = all digits argmax-correct
//   loss = sum(mask * per-token) / sum(mask)
// logits are fp32 [B*T*S, nv*vs] in (b, t, s) order (nv*vs <= 1024, vs % 128 == 0). One warp per
// token; each row is read once in forward (and once more in backward, which recomputes the
// softmax from the saved log-sum-exp instead of storing probabilities).
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct CeParams {
  const float* logits;
  long long ld;
  const long long* labels;     // [B*T*S]
  const long long* input_ids;  // [B*T*S]
  int B, T, S, nv, vs;
  long long mask_id;
  float smoothing;
  float* lse;    // [rows, nv]
  float* sums;   // [3]: sum(mask*loss), sum(mask*acc), sum(mask)
  // backward
  const float* dloss;  // device scalar
  __nv_bfloat16* dlogits;
  long long ldd;
};

__global__ void __launch_bounds__(256) ce_fwd_kernel(const CeParams p) {
  __shared__ float part[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)p.B * p.T * p.S;
  const long long row = (long long)blockIdx.x * 8 + warp;
  float loss = 0.f, acc = 0.f, cnt = 0.f;
  if (row < rows) {
    const int t = (int)((row / p.S) % p.T);
    const bool relevant = t >= 1 && p.input_ids[row] == p.mask_id;
    const long long label = p.labels[row];
    const float* z = p.logits + (size_t)row * p.ld;
    bool all_ok = true;
    long long div = 1;
    // rows outside the masked set contribute nothing (and get a zero gradient): skip them (warp-uniform)
    for (int k = 0; relevant && k < p.nv; ++k) {
      const int y = (int)((label / div) % p.vs);
      div *= p.vs;
      const float* zk = z + k * p.vs;
      float m = -INFINITY, zsum = 0.f;
      int am = 0;
      for (int c = lane * 4; c < p.vs; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(zk + c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zsum += vv[j];
          if (vv[j] > m) { m = vv[j]; am = c + j; }
        }
      }
      // warp arg-max with first-index tie break (torch.argmax semantics)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
      }
      zsum = warp_sum(zsum);
      float e = 0.f;
      for (int c = lane * 4; c < p.vs; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(zk + c);
        e += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
      }
      e = warp_sum(e);
      const float lse = m + logf(e);
      if (lane == 0 && p.lse != nullptr) p.lse[row * p.nv + k] = lse;
      const float zy = zk[y];
      loss += (1.0f - p.smoothing) * (lse - zy) + p.smoothing * (lse - zsum / (float)p.vs);
      all_ok = all_ok && (am == y);
    }
    if (relevant) { cnt = 1.f; acc = all_ok ? 1.f : 0.f; }
  }
  if (lane == 0) { part[warp][0] = loss; part[warp][1] = acc; part[warp][2] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
    if (s != 0.f) atomicAdd(p.sums + threadIdx.x, s);
  }
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const CeParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = (long long)p.B * p.T * p.S;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int t = (int)((row / p.S) % p.T);
  const bool relevant = t >= 1 && p.input_ids[row] == p.mask_id;
  __nv_bfloat16* d = p.dlogits + (size_t)row * p.ldd;
  const int width = p.nv * p.vs;
  if (!relevant) {
    for (int c = lane * 8; c < width; c += 256) *reinterpret_cast<uint4*>(d + c) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float coef = __ldg(p.dloss) / fmaxf(__ldg(p.sums + 2), 1.0f);
  const long long label = p.labels[row];
  const float* z = p.logits + (size_t)row * p.ld;
  const float eps_v = p.smoothing / (float)p.vs;
  long long div = 1;
  for (int k = 0; k < p.nv; ++k) {
    const int y = (int)((label / div) % p.vs);
    div *= p.vs;
    const float lse = p.lse[row * p.nv + k];
    for (int c = lane * 8; c < p.vs; c += 256) {
      const float4 a = *reinterpret_cast<const float4*>(z + k * p.vs + c);
      const float4 b = *reinterpret_cast<const float4*>(z + k * p.vs + c + 4);
      const float vv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = expf(vv[j] - lse) - eps_v;
        if (c + j == y) gj -= (1.0f - p.smoothing);
        g[j] = gj * coef;
      }
      *reinterpret_cast<uint4*>(d + k * p.vs + c) =
          make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
    }
  }
}

__global__ void ce_finalize_kernel(const float* sums, float* out) {
  // out[0] = loss, out[1] = acc (0/0 -> NaN, as the reference's division would give)
  out[0] = sums[0] / sums[2];
  out[1] = sums[1] / sums[2];
}

}  // namespace hma

extern "C" int hma_ce_fwd(const float* logits, long long ld, const long long* labels, const long long* input_ids,
                          int B, int T, int S, int nv, int vs, long long mask_id, float smoothing, float* lse,
                          float* sums, float* loss_acc, void* stream_) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HMA_REQUIRE(vs % 128 == 0 && nv >= 1 && nv * vs <= 1024, "ce_fwd: unsupported vocabulary %d x %d", nv, vs);
  HMA_REQUIRE(ld % 4 == 0, "ce_fwd: logits rows must be 16-byte aligned");
  CeParams p{};
  p.logits = logits; p.ld = ld; p.labels = labels; p.input_ids = input_ids;
  p.B = B; p.T = T; p.S = S; p.nv = nv; p.vs = vs; p.mask_id = mask_id; p.smoothing = smoothing;
  p.lse = lse; p.sums = sums;
  const long long rows = (long long)B * T * S;
  HMA_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), stream));
  if (rows > 0) {
    ce_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(p);
    HMA_CHECK_CUDA(cudaGetLastError());
  }
  ce_finalize_kernel<<<1, 1, 0, stream>>>(sums, loss_acc);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_ce_bwd(const float* logits, long long ld, const long long* labels, const long long* input_ids,
                          int B, int T, int S, int nv, int vs, long long mask_id, float smoothing, const float* lse,
                          const float* sums, const float* dloss, void* dlogits, long long ldd, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(vs % 256 == 0 && nv * vs <= 1024, "ce_bwd: unsupported vocabulary %d x %d", nv, vs);
  HMA_REQUIRE(ldd % 8 == 0 && ld % 4 == 0, "ce_bwd: rows must be 16-byte aligned");
  CeParams p{};
  p.logits = logits; p.ld = ld; p.labels = labels; p.input_ids = input_ids;
  p.B = B; p.T = T; p.S = S; p.nv = nv; p.vs = vs; p.mask_id = mask_id; p.smoothing = smoothing;
  p.lse = const_cast<float*>(lse); p.sums = const_cast<float*>(sums);
  p.dloss = dloss; p.dlogits = static_cast<__nv_bfloat16*>(dlogits); p.ldd = ldd;
  const long long rows = (long long)B * T * S;
  if (rows == 0) return 0;
  ce_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
