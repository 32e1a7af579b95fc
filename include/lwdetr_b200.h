/*
 * lwdetr_b200 - C ABI of the B200-native LW-DETR inference path.
 *
 * Every entry point takes plain pointers and sizes (no torch types).  Device pointers are CUDA
 * device addresses on the current device; `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  All calls are stream-ordered and asynchronous unless stated otherwise.
 * Return value: 0 on success, non-zero on failure with a message in lwdetr_last_error()
 * (thread-local).  Nothing throws across this boundary.
 *
 * The reference interface each group replaces is cited as file:line relative to the
 * Atten4Vis/LW-DETR tree.
 */
#ifndef LWDETR_B200_H_
#define LWDETR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LWDETR_API __attribute__((visibility("default")))

enum { LWDETR_F16 = 0, LWDETR_BF16 = 1 };
enum { LWDETR_ACT_NONE = 0, LWDETR_ACT_RELU = 1, LWDETR_ACT_GELU = 2, LWDETR_ACT_SILU = 3 };

LWDETR_API const char* lwdetr_last_error(void);
LWDETR_API int lwdetr_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level entry points (unit tests, ncu captures).  They are the same kernels the model-level
 * forward schedules.
 * ---------------------------------------------------------------------------------------------- */

/* out[rowmap(m), n] = resid[m', n] + gamma[n] * act(sum_k A[m,k] W[n,k] + bias[n])
 * tcgen05/TMA GEMM standing in for F.linear / 1x1 conv / ConvTranspose2d(k2,s2)
 * (models/backbone/vit.py:120-140,206-220; projector.py:85-98,165-190; transformer.py:27-39).
 *   A [M, lda] 16-bit, W [N, K] 16-bit, K % 64 == 0, bias/gamma fp32 [N] or NULL,
 *   resid 16-bit [*, ld_resid] or NULL (row m % resid_mod when resid_mod > 0),
 *   out 16-bit (out_fp32 = 0) or fp32 [*, ld_out].
 *   rows_in = 1: rows are window-major tokens (b, win 4x4, t) of an IH x IW grid (vit.py:353-358);
 *   remap_rows = 1: write rows in spatial (b, y, x) order (the un-windowing of vit.py:362-364);
 *   shuffle_cout > 0: N == 4*cout, column (dy*2+dx)*cout+co goes to pixel (2y+dy, 2x+dx), channel co. */
LWDETR_API int lwdetr_gemm(int dtype, const void* A, int lda, int M, int K, const void* W, int N,
                           const float* bias, const float* gamma, const void* resid, int ld_resid, int resid_mod,
                           int act, void* out, int ld_out, int out_fp32, int rows_in, int remap_rows,
                           int shuffle_cout, int IH, int IW, void* stream);

/* 3x3 convolution, padding 1, stride 1 or 2, NHWC, as an implicit GEMM (no im2col buffer):
 * X [B, stride*OH, stride*OW, ldx] (channel slice of width Cin), W [N, 9*Cin] with k = (dy*3+dx)*Cin + c,
 * out [B*OH*OW, ld_out].  Stands in for ConvX (projector.py:85-98) with BN folded into W/bias. */
LWDETR_API int lwdetr_conv3x3(int dtype, const void* X, int ldx, int B, int OH, int OW, int stride, int Cin,
                              const void* W, int N, const float* bias, int act, void* out, int ld_out,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LWDETR_B200_H_ */
