/*
 * hma_b200 — C ABI of the B200-native ST-MaskGIT hot path.
 *
 * The reference (liruiw/HMA) has no FFI: its hot path sits behind the Python nn.Module API of
 * hma/model/st_mask_git.py (STMaskGIT) and hma/model/st_transformer.py (STBlock). This header is
 * what a binding for that path would call; hma_b200/ (Python, ctypes) is such a binding and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated here
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised
 *   - return 0 on success, negative on error; hma_last_error() describes the last failure of
 *     the calling thread. No exceptions cross the boundary. There is NO CPU fallback.
 *   - activations/weights consumed by tensor-core kernels are bf16 (row-major, leading
 *     dimension in ELEMENTS); the residual stream, statistics, losses and gradients are fp32
 */
#ifndef HMA_B200_H_
#define HMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMA_B200_ABI_VERSION 1

int hma_abi_version(void);
const char* hma_last_error(void);
/* 0 iff the current CUDA device is sm_100 (B200). */
int hma_device_check(void);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core contractions (tcgen05 + TMEM + TMA)
 * ------------------------------------------------------------------------------------------- */

/* Epilogues of hma_gemm_nt */
#define HMA_EPI_BF16 0       /* out(bf16)  = alpha*acc + bias                                  */
#define HMA_EPI_GELU_BF16 1  /* z = alpha*acc + bias; out2(bf16) = z (optional); out = gelu(z) */
#define HMA_EPI_DGELU_BF16 2 /* out(bf16)  = (alpha*acc) * gelu'(aux)                          */
#define HMA_EPI_RESID_F32 3  /* out(fp32)  = resid(fp32, optional) + alpha*acc + bias          */

/* out[M,N] = epi(A[M,K] . B[N,K]^T). A, B bf16 row-major. Replaces nn.Linear forward
 * (attention.py:141,154; st_transformer.py:24-27; st_mask_git.py:70-75,681-683) and, with B a
 * pre-transposed weight, its input-gradient. K % 64 == 0, N % 128 == 0. */
int hma_gemm_nt(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epi,
                void* out, long long ldo, void* out2, long long ldo2, const float* bias, const float* resid,
                long long ldr, const void* aux, long long ldaux, float alpha, void* stream);

/* dW[Mw,Nw] (fp32) += G[tokens,Mw]^T . X[tokens,Nw]; G, X bf16 row-major. The weight gradient of
 * the same Linears. Mw % 128 == 0, Nw % 128 == 0. Accumulates (caller zeroes dW). */
int hma_gemm_wgrad(const void* G, long long ldg, const void* X, long long ldx, int tokens, int Mw, int Nw,
                   float* dW, long long ldw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMA_B200_H_ */
