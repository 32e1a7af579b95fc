// Multi-scale deformable attention forward (see msda.cu).
#pragma once
#include <cuda_runtime.h>
#include "gemm_tc.h"

namespace lwb {

static constexpr int MSDA_MAX_LEVELS = 4;

struct MsdaArgs {
  const void* value;        // 16-bit [B, S, ldv]; head m, channel c at column m*16 + c (of this layer's slice)
  int ldv;
  const void* offs_logits;  // 16-bit [B*nq, ld_ol]: [M*L*P*2 sampling offsets | M*L*P attention logits]
  int ld_ol;
  const float* ref;         // fp32 [B*nq, 4] reference boxes (cx, cy, w, h), un-sigmoided space
  void* out;                // 16-bit [B*nq, ld_out]
  int ld_out;
  int batch, nq, heads, levels, points, S;
  int lvl_h[MSDA_MAX_LEVELS], lvl_w[MSDA_MAX_LEVELS], lvl_start[MSDA_MAX_LEVELS];
};

int msda_launch(int dtype, const MsdaArgs& a, cudaStream_t st);

}  // namespace lwb
