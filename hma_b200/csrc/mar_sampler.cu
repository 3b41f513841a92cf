// The ancestral sampling loop of the diffusion head (DiffLoss.sample -> p_sample_loop: diffloss.py:37-59,
// gaussian_diffusion.py:394-490; network SimpleMLPAdaLN: diffloss.py:163-233) as ONE persistent kernel.
//
// Launched kernel by kernel, a spaced step of the sampler is 16 dependent launches over <= a few hundred rows: ~135 us of
// launch / set-up / drain latency for ~9 us of arithmetic (tools/ubench/sampler_chain.py: 7.1 us per 1024x1024 GEMM whatever
// its row count, 4.0 us per row-wise kernel). Here every CTA stays resident for ALL steps of a call and the stages of a step
// are separated by a grid-wide barrier (one atomic + an acquire spin, ~1.5 us) instead of a kernel boundary:
//
//   per step i (hi-1 .. lo), per residual block b:
//     GEMM1   a  = SiLU(u W1_b^T + b1_b)        128 x 32 output tiles over the grid: TMA -> tcgen05.mma -> TMEM -> epilogue
//     GEMM2   h2 = a W2_b^T + b2_b
//     ROW     x += gate_b * h2;  u = LN(x) * gamma_{b+1} + beta_{b+1}, modulated by (1 + scale_{b+1}) / shift_{b+1}
//   the ROW stage after the last block is the whole tail of the step in one pass over the row (a warp owns a row):
//     x += gate * h2;  uf = LN(x) (1 + scale_f) + shift_f;  (eps | v) = uf Wf^T + bf   (32 outputs: CUDA cores, Wf in smem)
//     x_{t-1} = p_sample(x_t, eps, v, noise_i)  and, for the next step, x = x_{t-1} W_in^T + b_in;  u = LN_0(x) modulated.
//
// The adaLN modulations of all steps are an input (one GEMM over steps x rows, mar.py); bf16 roundings are the ones of the
// kernel-by-kernel path (operands of every contraction bf16, accumulation and the residual stream fp32).
// All CTAs must be co-resident: the host launches at most one CTA per SM.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

constexpr int kSW = 1024;                     // width of the diffusion MLP
constexpr int kSBM = 128, kSBN = 32, kSBK = 64;
constexpr int kSStages = 10;                  // 200 KB in flight per CTA: a tile's 16 k-blocks are a latency chain otherwise
constexpr int kSAStage = kSBM * kSBK * 2;     // 16 KB
constexpr int kSBStage = kSBN * kSBK * 2;     // 4 KB
constexpr int kSStage = kSAStage + kSBStage;  // 20 KB (a multiple of 1024: swizzle atoms stay aligned)
constexpr int kSMaxDepth = 8;
constexpr int kSMaxD = 16;                    // token dimension (patch^2 x vae channels)
constexpr int kSThreads = 256;
constexpr int kSWfBytes = 2 * kSMaxD * kSW * 2;   // final linear [2D, 1024] bf16
constexpr int kSWinBytes = kSMaxD * kSW * 2;      // input projection, transposed [D, 1024] bf16
constexpr int kSSmem = 1024 + kSStages * kSStage;
static_assert(kSWfBytes + kSWinBytes <= kSStages * kSStage, "the narrow matrices borrow the (idle) ring during row stages");

struct SamplerMaps {
  CUtensorMap a_u, a_a;                       // A operands: u16, a16 [R, 1024]
  CUtensorMap w1[kSMaxDepth], w2[kSMaxDepth]; // B operands: mlp.0 / mlp.2 weights [1024, 1024]
};

struct SamplerParams {
  int R, D, depth;
  int step_hi, step_lo;                       // spaced steps step_hi-1 ... step_lo are processed
  float temperature;
  int clip;
  float* xt;                                  // [R, D], updated in place
  const float* noise;                         // [steps, R, D]
  const float* tables;                        // [steps, 8]
  const __nv_bfloat16* mods;                  // row (step - mods_step0) * R + r, columns: per block shift | scale | gate, then
  long long ldmod;                            // the final layer's shift | scale
  int mods_step0;
  const __nv_bfloat16* w_in_t;                // input projection transposed: [D, 1024]
  const float* b_in;
  const float* ln_g[kSMaxDepth];
  const float* ln_b[kSMaxDepth];
  const float* b1[kSMaxDepth];
  const float* b2[kSMaxDepth];
  const __nv_bfloat16* w_f;                   // [>= 2D, 1024]
  const float* b_f;
  float* x;                                   // [R, 1024] residual stream
  __nv_bfloat16* u16;
  __nv_bfloat16* a16;
  __nv_bfloat16* h2;
  float* dbg_out;                             // optional [R, 2D]: network output of the last step processed
  unsigned* barrier;                          // one word, zeroed by the host before the launch
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Grid-wide barrier (all CTAs resident). Orders generic-proxy global writes before later async-proxy (TMA) reads of any CTA.
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  fence_proxy_async_all();
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned target = epoch * gridDim.x;
    unsigned v, spins = 0;
    uint64_t t0 = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if ((++spins & 0x3fffu) == 0) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kMbarTimeoutNs) __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  uint32_t tphase = 0;  // accumulator hand-over parity (one tile at a time)
};

struct SamplerShared {
  uint64_t full[kSStages];
  uint64_t empty[kSStages];
  uint64_t tfull, tempty;
  uint64_t wbar;          // arrival of the two narrow matrices in the ring area
  uint32_t tmem_slot;
};

// One GEMM stage: out16[R, 1024] = act(A[R, 1024] . W[1024, 1024]^T + bias) in 128 x 32 tiles spread over the grid.
__device__ __forceinline__ void gemm_stage(const CUtensorMap* mA, const CUtensorMap* mB, const float* bias, bool act_silu,
                                           __nv_bfloat16* out16, int R, uint32_t ring, SamplerShared& sh, uint32_t tmem,
                                           PipeState& ps, int warp, int lane) {
  const int m_tiles = (R + kSBM - 1) / kSBM;
  const int total = m_tiles * (kSW / kSBN);
  constexpr int KB = kSW / kSBK;
  constexpr uint32_t kIdesc = umma_idesc_bf16(kSBM, kSBN, 0, 0);
  if (warp == 0) {
    if (lane == 0) {  // (always the same thread: the ring position lives in its registers across stages)
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (tile / (kSW / kSBN)) * kSBM, n0 = (tile % (kSW / kSBN)) * kSBN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.empty[ps.stage]), ps.phase ^ 1u);
          const uint32_t full = smem_u32(&sh.full[ps.stage]);
          mbar_expect_tx(full, (uint32_t)kSStage);
          tma_load_2d(ring + ps.stage * kSStage, mA, full, kb * kSBK, m0);
          tma_load_2d(ring + ps.stage * kSStage + kSAStage, mB, full, kb * kSBK, n0);
          if (++ps.stage == kSStages) { ps.stage = 0; ps.phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(smem_u32(&sh.tempty), ps.tphase ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.full[ps.stage]), ps.phase);
          tc_fence_after();
          const uint32_t a = ring + ps.stage * kSStage, b = a + kSAStage;
#pragma unroll
          for (int k = 0; k < kSBK / 16; ++k)
            umma_ss(tmem, umma_desc_kmajor(a + k * 32), umma_desc_kmajor(b + k * 32), kIdesc, (uint32_t)((kb | k) != 0));
          umma_commit(smem_u32(&sh.empty[ps.stage]));
          if (++ps.stage == kSStages) { ps.stage = 0; ps.phase ^= 1u; }
        }
        umma_commit(smem_u32(&sh.tfull));
        ps.tphase ^= 1u;
      }
    }
  } else if (warp >= 2 && warp < 6) {
    const int quarter = warp & 3;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m0 = (tile / (kSW / kSBN)) * kSBM, n0 = (tile % (kSW / kSBN)) * kSBN;
      mbar_wait(smem_u32(&sh.tfull), ps.tphase);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_x32(tmem_addr(tmem, (uint32_t)(quarter * 32), 0u), r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.tempty));
      ps.tphase ^= 1u;
      const int row = m0 + quarter * 32 + lane;
      if (row < R) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + j));
          float v0 = __uint_as_float(r[j]) + bb.x, v1 = __uint_as_float(r[j + 1]) + bb.y;
          float v2 = __uint_as_float(r[j + 2]) + bb.z, v3 = __uint_as_float(r[j + 3]) + bb.w;
          if (act_silu) { v0 = silu(v0); v1 = silu(v1); v2 = silu(v2); v3 = silu(v3); }
          pk[j >> 1] = pack_bf16(v0, v1);
          pk[(j >> 1) + 1] = pack_bf16(v2, v3);
        }
        uint4* dst = reinterpret_cast<uint4*>(out16 + (size_t)row * kSW + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
  }
}

// ---- row-wise pieces: a warp owns a row; lane l holds columns (32 k + l) * 4 .. + 3, k = 0..7.
// A row is one dependent chain per warp, so every global operand of a stage is requested BEFORE the first use of any of them
// (one memory latency per stage instead of one per operand group).
struct RowMod {   // LayerNorm affine (optional) and adaLN shift / scale of this lane's 32 columns
  float4 gm[8], bt[8];
  uint2 sh[8], sc[8];
};
__device__ __forceinline__ void row_mod_load(RowMod& m, const float* gamma, const float* beta, const __nv_bfloat16* modrow,
                                             int shift_off, int scale_off, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = (k * 32 + lane) * 4;
    m.sh[k] = *reinterpret_cast<const uint2*>(modrow + shift_off + col);
    m.sc[k] = *reinterpret_cast<const uint2*>(modrow + scale_off + col);
    if (gamma != nullptr) {
      m.gm[k] = __ldg(reinterpret_cast<const float4*>(gamma + col));
      m.bt[k] = __ldg(reinterpret_cast<const float4*>(beta + col));
    }
  }
}
__device__ __forceinline__ void row_ln_store(float4 (&v)[8], const RowMod& m, bool affine, __nv_bfloat16* urow, float (*keep)[4],
                                             int lane) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  const float mean = warp_sum(s) * (1.f / kSW);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
    q += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / kSW) + 1e-6f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = (k * 32 + lane) * 4;
    float o[4] = {v[k].x * rstd, v[k].y * rstd, v[k].z * rstd, v[k].w * rstd};
    if (affine) {
      o[0] = o[0] * m.gm[k].x + m.bt[k].x; o[1] = o[1] * m.gm[k].y + m.bt[k].y;
      o[2] = o[2] * m.gm[k].z + m.bt[k].z; o[3] = o[3] * m.gm[k].w + m.bt[k].w;
    }
    o[0] = o[0] * (1.f + bf16_lo(m.sc[k].x)) + bf16_lo(m.sh[k].x);
    o[1] = o[1] * (1.f + bf16_hi(m.sc[k].x)) + bf16_hi(m.sh[k].x);
    o[2] = o[2] * (1.f + bf16_lo(m.sc[k].y)) + bf16_lo(m.sh[k].y);
    o[3] = o[3] * (1.f + bf16_hi(m.sc[k].y)) + bf16_hi(m.sh[k].y);
    const uint2 pkd = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    if (urow != nullptr) *reinterpret_cast<uint2*>(urow + col) = pkd;
    if (keep != nullptr) {  // the bf16-rounded values, for a contraction done right here
      keep[k][0] = bf16_lo(pkd.x); keep[k][1] = bf16_hi(pkd.x); keep[k][2] = bf16_lo(pkd.y); keep[k][3] = bf16_hi(pkd.y);
    }
  }
}

// x (registers) += gate * h2
__device__ __forceinline__ void row_gate(float4 (&v)[8], const float* xrow, const __nv_bfloat16* modrow, int gate_off,
                                         const __nv_bfloat16* hrow, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = (k * 32 + lane) * 4;
    v[k] = *reinterpret_cast<const float4*>(xrow + col);
    const uint2 g = *reinterpret_cast<const uint2*>(modrow + gate_off + col);
    const uint2 h = *reinterpret_cast<const uint2*>(hrow + col);
    v[k].x = fmaf(bf16_lo(g.x), bf16_lo(h.x), v[k].x);
    v[k].y = fmaf(bf16_hi(g.x), bf16_hi(h.x), v[k].y);
    v[k].z = fmaf(bf16_lo(g.y), bf16_lo(h.y), v[k].z);
    v[k].w = fmaf(bf16_hi(g.y), bf16_hi(h.y), v[k].w);
  }
}

// x = bf16(x_t) W_in^T + b_in for this row (x_t value of element e in lane e), W_in^T [D, 1024] bf16 in shared memory
__device__ __forceinline__ void row_in_proj(float4 (&v)[8], float xt_lane, int D, const __nv_bfloat16* s_win, const float* b_in,
                                            int lane) {
  const float xb = __bfloat162float(__float2bfloat16(xt_lane));
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(b_in + (k * 32 + lane) * 4));
  for (int e = 0; e < D; ++e) {
    const float xe = __shfl_sync(0xffffffffu, xb, e);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint2 w = *reinterpret_cast<const uint2*>(s_win + (size_t)e * kSW + (k * 32 + lane) * 4);
      v[k].x = fmaf(xe, bf16_lo(w.x), v[k].x);
      v[k].y = fmaf(xe, bf16_hi(w.x), v[k].y);
      v[k].z = fmaf(xe, bf16_lo(w.y), v[k].z);
      v[k].w = fmaf(xe, bf16_hi(w.y), v[k].w);
    }
  }
}

__global__ void __launch_bounds__(kSThreads, 1) mar_sampler_kernel(const __grid_constant__ SamplerMaps maps, const SamplerParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ SamplerShared sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring = base;
  // The final linear Wf [2D, 1024] and the input projection W_in^T [D, 1024] are read by the row stages that end / begin a
  // step — when the TMA ring is idle: they are bulk-copied into its first 96 KB at the start of those stages.
  __nv_bfloat16* s_wf = reinterpret_cast<__nv_bfloat16*>(smem_raw + (base - smem_u32(smem_raw)));
  __nv_bfloat16* s_win = s_wf + 2 * kSMaxD * kSW;
  uint32_t wphase = 0;
  auto stage_narrow = [&]() {  // one thread; everybody waits on sh.wbar before touching s_wf / s_win
    const uint32_t bar = smem_u32(&sh.wbar);
    const uint32_t nf = (uint32_t)(2 * p.D * kSW * 2), ni = (uint32_t)(p.D * kSW * 2);
    mbar_expect_tx(bar, nf + ni);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base),
                 "l"(p.w_f), "r"(nf), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + kSWfBytes),
                 "l"(p.w_in_t), "r"(ni), "r"(bar) : "memory");
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_u);
    tma_prefetch_desc(&maps.a_a);
    for (int s = 0; s < kSStages; ++s) {
      mbar_init(smem_u32(&sh.full[s]), 1);
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    mbar_init(smem_u32(&sh.tfull), 1);
    mbar_init(smem_u32(&sh.tempty), 128);
    mbar_init(smem_u32(&sh.wbar), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&sh.tmem_slot), 32);
    tmem_relinquish();
  }
  pdl_wait();
  const int D = p.D, D2 = 2 * p.D;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_slot;
  if (threadIdx.x == 0) stage_narrow();

  PipeState ps;
  unsigned epoch = 0;
  const int gw = blockIdx.x * (kSThreads / 32) + warp, gws = gridDim.x * (kSThreads / 32);
  const int w3 = 3 * kSW;
  const int fin_off = w3 * p.depth;

  // ---- first stage of the call: x = in_proj(x_t), u = LN_0(x) modulated for step hi-1
  {
    const int step = p.step_hi - 1;
    mbar_wait(smem_u32(&sh.wbar), wphase);
    wphase ^= 1u;
    for (int r = gw; r < p.R; r += gws) {
      float4 v[8];
      const float xl = lane < D ? p.xt[(size_t)r * D + lane] : 0.f;
      row_in_proj(v, xl, D, s_win, p.b_in, lane);
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(p.x + (size_t)r * kSW + (k * 32 + lane) * 4) = v[k];
      const __nv_bfloat16* modrow = p.mods + ((size_t)(step - p.mods_step0) * p.R + r) * p.ldmod;
      RowMod m;
      row_mod_load(m, p.ln_g[0], p.ln_b[0], modrow, 0, kSW, lane);
      row_ln_store(v, m, true, p.u16 + (size_t)r * kSW, nullptr, lane);
    }
  }
  grid_sync(p.barrier, epoch);

  // development switches (bits 1-3 of `clip`, set through HMA_SAMPLER_DBG by hma_b200.ops; results are then meaningless): leave
  // out the GEMM stages / the row stages / the grid barriers, to time what remains (tools/ubench/sampler_call.py)
  const bool dbg_no_gemm = (p.clip & 2) != 0, dbg_no_row = (p.clip & 4) != 0, dbg_no_sync = (p.clip & 8) != 0;
  for (int step = p.step_hi - 1; step >= p.step_lo; --step) {
    for (int blk = 0; blk < p.depth; ++blk) {
      if (!dbg_no_gemm) gemm_stage(&maps.a_u, &maps.w1[blk], p.b1[blk], true, p.a16, p.R, ring, sh, tmem, ps, warp, lane);
      if (!dbg_no_sync) grid_sync(p.barrier, epoch); else __syncthreads();
      if (!dbg_no_gemm) gemm_stage(&maps.a_a, &maps.w2[blk], p.b2[blk], false, p.h2, p.R, ring, sh, tmem, ps, warp, lane);
      if (!dbg_no_sync) grid_sync(p.barrier, epoch); else __syncthreads();
      const bool last = blk == p.depth - 1;
      if (last) {  // the ring is idle (every MMA of the stage before has completed): bring the narrow matrices in
        if (threadIdx.x == 0) stage_narrow();
      }
      bool narrow_ready = false;
      for (int r = gw; r < (dbg_no_row ? 0 : p.R); r += gws) {
        const __nv_bfloat16* modrow = p.mods + ((size_t)(step - p.mods_step0) * p.R + r) * p.ldmod;
        float4 v[8];
        RowMod m;
        if (!last) row_mod_load(m, p.ln_g[blk + 1], p.ln_b[blk + 1], modrow, w3 * (blk + 1), w3 * (blk + 1) + kSW, lane);
        else row_mod_load(m, nullptr, nullptr, modrow, fin_off, fin_off + kSW, lane);
        // operands of the ancestral update (tail only), requested with everything else
        float tb2 = 0.f, tb3 = 0.f, tb4 = 0.f, tb5 = 0.f, tb6 = 0.f, tb7 = 0.f, xv = 0.f, nz = 0.f, bfv = 0.f;
        if (last) {
          const float* tb = p.tables + (size_t)step * 8;
          tb2 = __ldg(tb + 2); tb3 = __ldg(tb + 3); tb4 = __ldg(tb + 4); tb5 = __ldg(tb + 5); tb6 = __ldg(tb + 6); tb7 = __ldg(tb + 7);
          if (lane < D) {
            xv = p.xt[(size_t)r * D + lane];
            nz = __ldg(p.noise + ((size_t)step * p.R + r) * D + lane);
          }
          if (lane < D2) bfv = __ldg(p.b_f + lane);
        }
        row_gate(v, p.x + (size_t)r * kSW, modrow, w3 * blk + 2 * kSW, p.h2 + (size_t)r * kSW, lane);
        if (!last) {
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(p.x + (size_t)r * kSW + (k * 32 + lane) * 4) = v[k];
          row_ln_store(v, m, true, p.u16 + (size_t)r * kSW, nullptr, lane);
          continue;
        }
        // ---- tail of the step: final layer, (eps | v), ancestral update, and the head of the next step
        float uf[8][4];
        row_ln_store(v, m, false, nullptr, uf, lane);
        if (!narrow_ready) {
          mbar_wait(smem_u32(&sh.wbar), wphase);
          narrow_ready = true;
        }
        if (step > p.step_lo)  // LN_0 operands of the next step: in flight during the output projection
          row_mod_load(m, p.ln_g[0], p.ln_b[0], p.mods + ((size_t)(step - 1 - p.mods_step0) * p.R + r) * p.ldmod, 0, kSW, lane);
        float acc[2 * kSMaxD];
#pragma unroll
        for (int o = 0; o < 2 * kSMaxD; ++o) {
          acc[o] = 0.f;
          if (o < D2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint2 w = *reinterpret_cast<const uint2*>(s_wf + (size_t)o * kSW + (k * 32 + lane) * 4);
              acc[o] = fmaf(uf[k][0], bf16_lo(w.x), acc[o]);
              acc[o] = fmaf(uf[k][1], bf16_hi(w.x), acc[o]);
              acc[o] = fmaf(uf[k][2], bf16_lo(w.y), acc[o]);
              acc[o] = fmaf(uf[k][3], bf16_hi(w.y), acc[o]);
            }
          }
        }
        // butterfly: lane l ends with the sum over lanes of acc[l]
#pragma unroll
        for (int s = kSMaxD; s >= 1; s >>= 1) {
#pragma unroll
          for (int j = 0; j < s; ++j) {
            const bool up = (lane & s) != 0;
            const float keepv = up ? acc[j + s] : acc[j];
            const float send = up ? acc[j] : acc[j + s];
            acc[j] = keepv + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        const float outv = acc[0] + bfv;  // lane o: output o (eps: o < D, v: D <= o < 2D)
        if (p.dbg_out != nullptr && lane < D2) p.dbg_out[(size_t)r * D2 + lane] = outv;
        const float vv = __shfl_sync(0xffffffffu, outv, (lane + D) & 31);     // lane e < D gets v_e from lane D + e
        float nx = 0.f;
        if (lane < D) {
          const float frac = 0.5f * (vv + 1.f);
          const float lv = frac * tb7 + (1.f - frac) * tb6;
          float px0 = tb2 * xv - tb3 * outv;
          if (p.clip & 1) px0 = fminf(fmaxf(px0, -10.f), 10.f);
          const float mean = tb4 * px0 + tb5 * xv;
          nx = mean + (step != 0 ? expf(0.5f * lv) * nz * p.temperature : 0.f);
          p.xt[(size_t)r * D + lane] = nx;
        }
        if (step > p.step_lo) {
          row_in_proj(v, nx, D, s_win, p.b_in, lane);
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(p.x + (size_t)r * kSW + (k * 32 + lane) * 4) = v[k];
          row_ln_store(v, m, true, p.u16 + (size_t)r * kSW, nullptr, lane);
        }
      }
      if (last) {  // warps without a row also consume the phase, so that every thread's parity stays in step
        if (!narrow_ready) mbar_wait(smem_u32(&sh.wbar), wphase);
        wphase ^= 1u;
      }
      if (!dbg_no_sync) grid_sync(p.barrier, epoch); else __syncthreads();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace hma

// The ancestral steps step_hi-1 ... step_lo of the diffusion head's sampler for R rows in one launch (see the header comment).
// w_in_t: the input projection TRANSPOSED, bf16 [D, 1024] contiguous. w1 / w2: `depth` device pointers each (bf16 [1024, 1024], row stride 1024); ln_g / ln_b / b1 / b2: `depth` fp32 [1024]
// pointers each (host arrays). Workspaces x (fp32), u16 / a16 / h2 (bf16): [R, 1024]; barrier: one u32 (zeroed here).
extern "C" int hma_mar_sampler(int R, int D, int depth, int step_hi, int step_lo, float temperature, int clip, float* xt,
                               const float* noise, const float* tables, const void* mods, long long ldmod, int mods_step0,
                               const void* w_in_t, const float* b_in, const void* const* w1,
                               const void* const* w2, const float* const* ln_g, const float* const* ln_b,
                               const float* const* b1, const float* const* b2, const void* w_f, const float* b_f, float* x,
                               void* u16, void* a16, void* h2, float* dbg_out, unsigned* barrier, void* stream_) {
  using namespace hma;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (R == 0 || step_hi <= step_lo) return 0;
  HMA_REQUIRE(D >= 1 && D <= kSMaxD, "mar_sampler: token dimension %d not in [1, %d]", D, kSMaxD);
  HMA_REQUIRE(depth >= 1 && depth <= kSMaxDepth, "mar_sampler: depth %d not in [1, %d]", depth, kSMaxDepth);
  HMA_REQUIRE(step_lo >= mods_step0, "mar_sampler: modulations start at step %d, asked for step %d", mods_step0, step_lo);
  HMA_REQUIRE(ldmod % 4 == 0, "mar_sampler: modulation rows must be 8-byte aligned");
  SamplerMaps maps;
  int rc = hma_host::make_tmap_bf16_2d(&maps.a_u, u16, kSW, (uint64_t)R, (uint64_t)kSW * 2, kSBK, kSBM);
  if (rc) return rc;
  rc = hma_host::make_tmap_bf16_2d(&maps.a_a, a16, kSW, (uint64_t)R, (uint64_t)kSW * 2, kSBK, kSBM);
  if (rc) return rc;
  SamplerParams p;
  for (int i = 0; i < depth; ++i) {
    rc = hma_host::make_tmap_bf16_2d(&maps.w1[i], w1[i], kSW, kSW, (uint64_t)kSW * 2, kSBK, kSBN);
    if (rc) return rc;
    rc = hma_host::make_tmap_bf16_2d(&maps.w2[i], w2[i], kSW, kSW, (uint64_t)kSW * 2, kSBK, kSBN);
    if (rc) return rc;
    p.ln_g[i] = ln_g[i]; p.ln_b[i] = ln_b[i]; p.b1[i] = b1[i]; p.b2[i] = b2[i];
  }
  for (int i = depth; i < kSMaxDepth; ++i) {
    maps.w1[i] = maps.w1[0]; maps.w2[i] = maps.w2[0];
    p.ln_g[i] = p.ln_b[i] = p.b1[i] = p.b2[i] = nullptr;
  }
  p.R = R; p.D = D; p.depth = depth; p.step_hi = step_hi; p.step_lo = step_lo;
  p.temperature = temperature; p.clip = clip;
  p.xt = xt; p.noise = noise; p.tables = tables;
  p.mods = static_cast<const __nv_bfloat16*>(mods); p.ldmod = ldmod; p.mods_step0 = mods_step0;
  p.w_in_t = static_cast<const __nv_bfloat16*>(w_in_t); p.b_in = b_in;
  p.w_f = static_cast<const __nv_bfloat16*>(w_f); p.b_f = b_f;
  p.x = x; p.u16 = static_cast<__nv_bfloat16*>(u16); p.a16 = static_cast<__nv_bfloat16*>(a16);
  p.h2 = static_cast<__nv_bfloat16*>(h2);
  p.dbg_out = dbg_out; p.barrier = barrier;
  static hma_host::PerDeviceFlag attr_flag;
  bool& attr_done = attr_flag.get();
  if (!attr_done) {
    HMA_CHECK_CUDA(cudaFuncSetAttribute(mar_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSSmem));
    attr_done = true;
  }
  const int m_tiles = (R + kSBM - 1) / kSBM;
  int grid = m_tiles * (kSW / kSBN);
  const int row_ctas = (R + 7) / 8;
  if (grid < row_ctas) grid = row_ctas;
  if (grid > hma_host::sm_count()) grid = hma_host::sm_count();  // every CTA must be resident: the stages meet at a grid barrier
  HMA_CHECK_CUDA(cudaMemsetAsync(barrier, 0, sizeof(unsigned), stream));
  HMA_CHECK_CUDA(hma_host::launch_pdl(mar_sampler_kernel, dim3(grid), dim3(kSThreads), (size_t)kSSmem, stream, maps, p));
  return 0;
}
