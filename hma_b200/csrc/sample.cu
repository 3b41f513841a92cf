// MaskGIT sampling step, entirely on the device.
// Reference: st_mask_git.py:397-420 (per-vocabulary softmax, greedy argmax or Categorical sample,
// id = hi*vs + lo, confidence = product of the chosen probabilities) and :422-453 (cosine-schedule
// re-masking: rank the keys, re-mask the n smallest, unmask the rest, restore tokens that were
// already unmasked, write the frame back into the prompt in place).
//
// Random numbers are NOT drawn here: the host binding materialises exactly the tensors the
// reference draws (an Exp(1) tensor per vocabulary half, high half first,
// then a U(0,1) tensor for the random unmask order) with torch's own generator, so the Philox
// stream is consumed identically; these kernels do the deterministic part.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct SampleParams {
  const float* logits;       // token (b, s) at logits + b*stride_b + s*ld
  long long stride_b, ld;
  int B, S, nv, vs;
  const float* exp_noise;    // [nv][B*S, vs], index 0 = HIGH half (reference order) or null = greedy
  float temperature;         // divides the probabilities before the renormalisation (cancels in exact arithmetic)
  long long* samples;        // [B*S]
  float* conf;               // [B*S]
};

// one warp per token
__global__ void __launch_bounds__(256) sample_tokens_kernel(const SampleParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * 8 + warp;
  if (tok >= (long long)p.B * p.S) return;
  const int b = (int)(tok / p.S), s = (int)(tok % p.S);
  const float* z = p.logits + (size_t)b * p.stride_b + (size_t)s * p.ld;
  long long id = 0;
  float conf = 1.f;
  for (int j = 0; j < p.nv; ++j) {
    const int k = p.nv - 1 - j;  // high factor first (flip(2), st_mask_git.py:408)
    const float* zk = z + k * p.vs;
    float m = -INFINITY;
    for (int c = lane * 4; c < p.vs; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(zk + c);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    m = warp_max(m);
    // torch.softmax: p = exp(z - m) / sum. The reference then ranks Categorical(probs = p / temperature), i.e.
    // ((p / temperature) / sum(p / temperature)) / Exp(1) (st_mask_git.py:409-416; torch.multinomial = argmax(probs / q)).
    // The keys below are built with the same sequence of fp32 operations, so near-ties round the way the reference's do.
    float sum = 0.f;
    for (int c = lane * 4; c < p.vs; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(zk + c);
      sum += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
    }
    sum = warp_sum(sum);
    const float* q = p.exp_noise != nullptr ? p.exp_noise + ((size_t)j * p.B * p.S + tok) * p.vs : nullptr;
    float sum2 = 1.f;
    if (q != nullptr) {
      sum2 = 0.f;
      for (int c = lane * 4; c < p.vs; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(zk + c);
        sum2 += (expf(v.x - m) / sum) / p.temperature + (expf(v.y - m) / sum) / p.temperature +
                (expf(v.z - m) / sum) / p.temperature + (expf(v.w - m) / sum) / p.temperature;
      }
      sum2 = warp_sum(sum2);
    }
    float best = -INFINITY, best_p = 0.f;
    int arg = 0;
    for (int c = lane * 4; c < p.vs; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(zk + c);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      float qq[4] = {1.f, 1.f, 1.f, 1.f};
      if (q != nullptr) {
        const float4 t4 = *reinterpret_cast<const float4*>(q + c);
        qq[0] = t4.x; qq[1] = t4.y; qq[2] = t4.z; qq[3] = t4.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pr = expf(vv[i] - m) / sum;
        const float key = q != nullptr ? ((pr / p.temperature) / sum2) / qq[i] : pr;  // greedy: argmax of the probabilities
        if (key > best) { best = key; best_p = pr; arg = c + i; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const float oe = __shfl_xor_sync(0xffffffffu, best_p, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; best_p = oe; arg = oa; }
    }
    id = id * p.vs + arg;
    conf *= best_p;
  }
  if (lane == 0) {
    p.samples[tok] = id;
    p.conf[tok] = conf;
  }
}

struct RemaskParams {
  const float* keys;        // [B, S] confidences or uniform noise (unused when n_mask < 0)
  unsigned char* unmasked;  // [B, S] in/out
  const long long* samples; // [B, S] freshly sampled ids
  long long* frame;         // prompt[:, out_t] flattened: token (b, s) at frame + b*stride_b + s; in/out
  long long stride_b;
  int B, S;
  int n_mask;               // tokens to re-mask; < 0 on the last step (no ranking)
  long long mask_id;
  long long* out_samples;   // [B, S] final samples of this step (what the reference returns)
};

// one CTA per sample, one thread per token (S <= 1024)
__global__ void __launch_bounds__(1024) rank_remask_kernel(const RemaskParams p) {
  extern __shared__ float skey[];
  const int b = blockIdx.x;
  const int i = threadIdx.x;
  const bool valid = i < p.S;
  bool was_unmasked = false;
  float key = INFINITY;
  if (valid) {
    was_unmasked = p.unmasked[(size_t)b * p.S + i] != 0;
    if (p.n_mask >= 0 && !was_unmasked) key = p.keys[(size_t)b * p.S + i];
    skey[i] = key;
  }
  __syncthreads();
  if (!valid) return;
  long long out = p.samples[(size_t)b * p.S + i];
  if (p.n_mask >= 0) {
    // rank in ascending (key, index) order == position in a stable argsort
    int rank = 0;
    for (int j = 0; j < p.S; ++j) {
      const float kj = skey[j];
      rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
    }
    if (rank < p.n_mask) out = p.mask_id;
    else p.unmasked[(size_t)b * p.S + i] = 1;
  }
  long long* slot = p.frame + (size_t)b * p.stride_b + i;
  if (was_unmasked) out = *slot;  // keep tokens fixed in earlier steps (st_mask_git.py:449)
  *slot = out;                    // in-place write into the caller's prompt (:453)
  p.out_samples[(size_t)b * p.S + i] = out;
}

}  // namespace hma

extern "C" int hma_sample_tokens(const float* logits, long long stride_b, long long ld, int B, int S, int nv, int vs,
                                 const float* exp_noise, float temperature, long long* samples, float* conf,
                                 void* stream_) {
  using namespace hma;
  HMA_REQUIRE(vs % 128 == 0 && nv >= 1 && nv <= 3, "sample_tokens: unsupported vocabulary %d x %d", nv, vs);
  HMA_REQUIRE(ld % 4 == 0 && stride_b % 4 == 0, "sample_tokens: logits must be 16-byte aligned");
  if (B * S == 0) return 0;
  HMA_REQUIRE(exp_noise == nullptr || temperature > 0.f, "sample_tokens: sampling needs temperature > 0");
  SampleParams p{logits, stride_b, ld, B, S, nv, vs, exp_noise, temperature, samples, conf};
  sample_tokens_kernel<<<(B * S + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hma_rank_remask(const float* keys, unsigned char* unmasked, const long long* samples, long long* frame,
                               long long stride_b, int B, int S, int n_mask, long long mask_id,
                               long long* out_samples, void* stream_) {
  using namespace hma;
  HMA_REQUIRE(S >= 1 && S <= 1024, "rank_remask: S=%d must be in [1,1024]", S);
  HMA_REQUIRE(n_mask < 0 || keys != nullptr, "rank_remask: ranking requested without keys");
  if (B == 0) return 0;
  RemaskParams p{keys, unmasked, samples, frame, stride_b, B, S, n_mask, mask_id, out_samples};
  const int threads = (S + 31) / 32 * 32;
  rank_remask_kernel<<<B, threads, S * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(p);
  HMA_CHECK_CUDA(cudaGetLastError());
  return 0;
}
