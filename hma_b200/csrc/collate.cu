// On-device MaskGIT training collator (reference: hma/data.py:28-98, get_maskgit_collator): random token corruption
// of the factorised ids (Copilot-4D style), the "non-MLM" branch that corrupts later frames progressively, and the
// per-(sample, frame) cosine-rate masking, in one pass over the tokens. The random draws are made by the caller (torch's
// generator, same tensors and order as the reference) and passed in: this kernel is the integer arithmetic on them,
// bit-exact against the reference given the same draws. One thread per token; every byte is streamed once.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

struct CollateParams {
  const long long* tokens;     // [B, T, S]
  long long* input_ids;        // [B, T, S]
  long long* labels;           // [B, T, S]
  int B, T, S, nv, vs;
  long long mask_id;
  const float* corrupt_r;      // [B, T, S, nv] or null
  float corrupt_thresh;        // max_corrupt_rate * u01
  const long long* rand_vals;  // [B, T, S, nv] (required by both corruption branches)
  int first_masked_frame;
  const float* frame_rates;    // [T - fmf] float32 or null (non-MLM branch)
  const float* frame_r;        // [B, T - fmf, S, nv]
  const float* mask_prob;      // [B, T - fmf] or null
  const float* mask_r;         // [B, T - fmf, S]
};

__global__ void __launch_bounds__(256) collate_maskgit_kernel(const CollateParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)p.B * p.T * p.S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % p.S);
    const long long bt = i / p.S;
    const int t = (int)(bt % p.T);
    const int b = (int)(bt / p.T);
    const long long tok = p.tokens[i];
    p.labels[i] = tok;
    long long out = 0, power = 1;
    const int tf = t - p.first_masked_frame;  // index into the per-masked-frame arrays
    const int Tm = p.T - p.first_masked_frame;
    for (int k = 0; k < p.nv; ++k) {
      long long f = (tok / power) % p.vs;     // factorize_token_ids (factorization_utils.py:57-68)
      if (p.corrupt_r != nullptr && p.corrupt_r[i * p.nv + k] < p.corrupt_thresh) f = p.rand_vals[i * p.nv + k];
      if (p.frame_r != nullptr && tf >= 0) {
        const float r = p.frame_r[(((long long)b * Tm + tf) * p.S + s) * p.nv + k];
        if (r > p.frame_rates[tf]) f = p.rand_vals[i * p.nv + k];
      }
      out += f * power;                       // unfactorize_token_ids
      power *= p.vs;
    }
    if (p.mask_r != nullptr && tf >= 0) {
      if (p.mask_r[((long long)b * Tm + tf) * p.S + s] < p.mask_prob[(long long)b * Tm + tf]) out = p.mask_id;
    }
    p.input_ids[i] = out;
  }
}

}  // namespace hma

extern "C" int hma_collate_maskgit(const long long* tokens, long long* input_ids, long long* labels, int B, int T, int S,
                                   int nv, int vs, long long mask_id, const float* corrupt_r, float corrupt_thresh,
                                   const long long* rand_vals, int first_masked_frame, const float* frame_rates,
                                   const float* frame_r, const float* mask_prob, const float* mask_r, void* stream_) {
  using namespace hma;
  if (B == 0) return 0;
  HMA_REQUIRE(B > 0 && T > 0 && S > 0 && nv >= 1 && vs >= 2, "collate_maskgit: bad shape");
  HMA_REQUIRE(first_masked_frame >= 0 && first_masked_frame <= T, "collate_maskgit: bad first_masked_frame %d", first_masked_frame);
  HMA_REQUIRE((corrupt_r == nullptr && frame_r == nullptr) || rand_vals != nullptr, "collate_maskgit: corruption needs rand_vals");
  HMA_REQUIRE((frame_r == nullptr) == (frame_rates == nullptr), "collate_maskgit: frame_r and frame_rates go together");
  HMA_REQUIRE((mask_r == nullptr) == (mask_prob == nullptr), "collate_maskgit: mask_r and mask_prob go together");
  CollateParams p{tokens, input_ids, labels, B, T, S, nv, vs, mask_id, corrupt_r, corrupt_thresh, rand_vals, first_masked_frame,
                  frame_rates, frame_r, mask_prob, mask_r};
  const long long total = (long long)B * T * S;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)hma_host::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  HMA_CHECK_CUDA(hma_host::launch_pdl(collate_maskgit_kernel, dim3((int)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream_), p));
  return 0;
}
