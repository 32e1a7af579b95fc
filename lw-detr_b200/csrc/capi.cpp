// C-ABI surface (include/lwdetr_b200.h): argument checking, error strings, dispatch to the kernels.
#include "lwdetr_b200.h"

#include <cuda_runtime.h>

#include <string>

#include "gemm_tc.h"

namespace {
thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return -1;
}
int cuda_fail(int e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(static_cast<cudaError_t>(e));
  return -1;
}
}  // namespace

extern "C" {

const char* lwdetr_last_error(void) { return g_err.c_str(); }
int lwdetr_abi_version(void) { return 1; }

int lwdetr_gemm(int dtype, const void* A, int lda, int M, int K, const void* W, int N, const float* bias,
                const float* gamma, const void* resid, int ld_resid, int resid_mod, int act, void* out, int ld_out,
                int out_fp32, int rows_in, int remap_rows, int shuffle_cout, int IH, int IW, void* stream) {
  if (!A || !W || !out) return fail("lwdetr_gemm: null pointer");
  lwb::GemmDesc d;
  d.dtype = dtype; d.A = A; d.lda = lda; d.M = M; d.N = N; d.K = K; d.W = W; d.bias = bias; d.gamma = gamma;
  d.resid = resid; d.ld_resid = ld_resid; d.resid_mod = resid_mod; d.act = act; d.out = out; d.ld_out = ld_out;
  d.out_fp32 = out_fp32; d.rows_in = rows_in; d.remap_rows = remap_rows; d.shuffle_cout = shuffle_cout;
  d.IH = IH; d.IW = IW;
  lwb::GemmOp op;
  std::string err;
  if (lwb::gemm_build(d, &op, &err)) return fail("lwdetr_gemm: " + err);
  int e = lwb::gemm_launch(op, static_cast<cudaStream_t>(stream));
  if (e) return cuda_fail(e, "lwdetr_gemm launch");
  return 0;
}

int lwdetr_conv3x3(int dtype, const void* X, int ldx, int B, int OH, int OW, int stride, int Cin, const void* W,
                   int N, const float* bias, int act, void* out, int ld_out, void* stream) {
  if (!X || !W || !out) return fail("lwdetr_conv3x3: null pointer");
  if (stride != 1 && stride != 2) return fail("lwdetr_conv3x3: stride must be 1 or 2");
  lwb::GemmDesc d;
  d.dtype = dtype; d.a_mode = stride == 1 ? lwb::AMODE_CONV3_S1 : lwb::AMODE_CONV3_S2;
  d.A = X; d.lda = ldx; d.B = B; d.OH = OH; d.OW = OW; d.M = B * OH * OW; d.N = N; d.K = 9 * Cin; d.W = W;
  d.bias = bias; d.act = act; d.out = out; d.ld_out = ld_out;
  lwb::GemmOp op;
  std::string err;
  if (lwb::gemm_build(d, &op, &err)) return fail("lwdetr_conv3x3: " + err);
  int e = lwb::gemm_launch(op, static_cast<cudaStream_t>(stream));
  if (e) return cuda_fail(e, "lwdetr_conv3x3 launch");
  return 0;
}

}  // extern "C"
