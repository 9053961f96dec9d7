/*
 * liboasr_b200 -- C ABI of the B200 (sm_100a) kernels behind olmoasr_b200.
 *
 * The reference (allenai/OLMoASR) has no FFI layer: its hot path is plain PyTorch library calls
 * made from olmoasr/model.py, olmoasr/inf_model.py and (third-party) whisper/audio.py.  Each entry
 * point below names the reference call it replaces (file:line under /root/reference).  The host side
 * (olmoasr_b200/*.py) binds these with ctypes; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - no allocation, no synchronisation, no global state besides a cached device-property query;
 *   - return 0 on success, a negative OASR_ERR_* otherwise; oasr_last_error() gives the message;
 *   - bf16 tensors are row-major, 16-byte aligned, row strides multiples of 8 elements unless noted.
 */
#ifndef OASR_B200_H_
#define OASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OASR_API __attribute__((visibility("default")))
#else
#define OASR_API
#endif

#define OASR_OK 0
#define OASR_ERR_INVALID (-1) /* bad argument (shape, alignment, enum)      */
#define OASR_ERR_CUDA (-2)    /* a CUDA runtime / driver call failed        */
#define OASR_ERR_UNSUPPORTED (-3)

/* ---- library ------------------------------------------------------------------------------ */
OASR_API const char* oasr_last_error(void); /* thread-local, valid until the next failing call          */
OASR_API int oasr_abi_version(void);
OASR_API int oasr_device_sm_count(void);

/* ---- dense GEMM on tcgen05 / TMEM ---------------------------------------------------------------
 * Replaces F.linear in Linear.forward (olmoasr/model.py:97-101), its autograd dgrad / wgrad, the tied
 * logits matmul (model.py:768-770) and the two Conv1d calls (model.py:592-593) after im2col.
 *
 *   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )        bf16 x bf16 -> fp32 accumulate
 *
 * Operand storage (the reduction index is k):
 *   OASR_K_MAJOR  : A stored [M][K] (lda = row stride), B stored [N][K]   -- k contiguous
 *   OASR_MN_MAJOR : A stored [K][M] (lda = row stride), B stored [K][N]   -- m / n contiguous
 * so forward y = x W^T is (K,K); dgrad dx = dy W is (K,MN); wgrad dW = dy^T x is (MN,MN).
 */
#define OASR_K_MAJOR 0
#define OASR_MN_MAJOR 1

#define OASR_EPI_BF16 0          /* C = bf16(acc + bf16(bias))                                  */
#define OASR_EPI_BF16_GELU 1     /* C = bf16(acc + bias) ; C2 = bf16(gelu_erf(C))               */
#define OASR_EPI_BF16_RESIDUAL 2 /* C = bf16(aux + bf16(acc + bias))                            */
#define OASR_EPI_BF16_GELU_BWD 3 /* C = bf16(bf16(acc) * gelu_erf'(aux))   aux = pre-activation  */
#define OASR_EPI_F32 4           /* C(f32) = acc + bias                                         */
#define OASR_EPI_F32_ATOMIC_ADD 5 /* C(f32) += acc   (split_k >= 1, C pre-initialised)           */

OASR_API int oasr_gemm_bf16(const void* A, int64_t lda, int a_layout, const void* B, int64_t ldb, int b_layout,
                   void* C, int64_t ldc, void* C2, const float* bias, const void* aux, int64_t ldaux,
                   int64_t M, int64_t N, int64_t K, int epilogue, int split_k, int block_n,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OASR_B200_H_ */
