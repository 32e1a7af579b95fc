// Fused attention core (see attn.cu).
#pragma once
#include <cuda_runtime.h>
#include "gemm_tc.h"   // DT_* codes

namespace lwb {

struct AttnArgs {
  const void* q; const void* k; const void* v;   // 16-bit row-major, head h at columns [h*dh, (h+1)*dh)
  int ldq, ldk, ldv;
  void* o; int ldo;
  int seqlen;        // tokens per sequence (100 window / 1600 global / nq decoder)
  int nseq;          // number of sequences; token (s, t) is matrix row s*seqlen + t
  int heads;
  float scale_log2;  // softmax scale * log2(e)
};

int attention_launch(int dtype, const AttnArgs& a, int dh, cudaStream_t st);

}  // namespace lwb
