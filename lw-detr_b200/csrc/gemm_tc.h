// tcgen05 + TMA GEMM family used for every linear / convolution of the LW-DETR forward.
//   out[row_map(m), n] = epilogue( sum_k A[m, k] * W[n, k] )
// A is gathered by TMA either from a plain row-major matrix, or implicitly (im2col-free) from an
// NHWC feature map for 3x3 convolutions (stride 1 and 2).  W is always [N, K] row-major, i.e. the
// nn.Linear / reordered conv weight layout, so both operands are K-major for the tensor core.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>

namespace lwb {

enum : int { DT_F16 = 0, DT_BF16 = 1 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3 };
enum : int { AMODE_PLAIN = 0, AMODE_CONV3_S1 = 1, AMODE_CONV3_S2 = 2 };
enum : int { ROWS_PLAIN = 0, ROWS_WINDOW_MAJOR = 1 };

// Device-visible arguments (passed by value to the kernel).
struct GemmArgs {
  int M, N, kblocks;        // valid rows / valid output columns / K in units of 64
  int n_tiles, m_tiles, stages;
  int a_mode;
  uint32_t a_stage_tx;      // bytes TMA delivers into the A slot per stage
  // conv geometry (output grid OH x OW per image, tile TW x TH pixels, Cin/64 channel blocks)
  int OH, OW, TW, TH, tiles_x, tiles_y, cin_blocks, lda;
  // epilogue
  const float* bias;        // [N] fp32 or null
  const float* gamma;       // [N] fp32 or null (layer-scale applied before the residual add)
  const void* resid;        // 16-bit [*, ld_resid] or null
  int ld_resid, resid_mod;  // residual row = m % resid_mod when resid_mod > 0
  int act;
  void* out;
  int ld_out, out_fp32;
  int rows_in;              // how input rows m are ordered: plain (b,y,x) or window-major (b,win,t)
  int remap_rows;           // 1: write rows in spatial (b,y,x) order (needs rows_in grid IH x IW)
  int shuffle_cout;         // >0: ConvTranspose2d(k=2,s=2) pixel shuffle, n = (dy*2+dx)*cout + co
  int IH, IW;               // token grid per image for rows_in / shuffle (40 x 40 for the ViT)
  // head-major store (the deformable-attention value tensor, msda.cu): hm_S > 0 writes element (row b*hm_S + s,
  // column slice*hm_heads*16 + head*16 + c) to out[(((b*hm_slices + slice)*hm_heads + head)*hm_S + s)*16 + c]
  int hm_S, hm_heads, hm_slices;
  const uint8_t* row_zero;  // [M] or null: rows flagged non-zero are stored as zeros (padded memory tokens, ms_deform_attn.py:114-115)
  // LayerNorm fusion (ViT blocks).  Producer side: the epilogue also emits per-row partial (sum, sum of
  // squares) of the rounded output, one float2 per (row, n_tile, column half).  Consumer side: the GEMM runs
  // on the RAW rows and applies  out = rstd*(acc - mean*colsum[n]) + bias[n]  where the LN weight is folded into
  // W, colsum[n] = sum_k W'[n,k] and bias already contains sum_k ln_b[k]*W[n,k].
  float2* stats_out;        // [M, stats_parts_out] or null
  int stats_parts_out;
  const float2* stats_in;   // [M, stats_parts_in] or null
  int stats_parts_in;
  const float* colsum;      // [N] fp32 (consumer)
  float ln_inv_c, ln_eps;   // 1/C of the normalised dimension, epsilon
  int epi_groups;           // 2: the two 8-warp epilogue groups take alternate tiles (many tiles per CTA); 1: both work on every tile
};

// Host description of one GEMM.
struct GemmDesc {
  int dtype = DT_F16;
  int a_mode = AMODE_PLAIN;
  const void* A = nullptr;  // plain: [M, lda]; conv: NHWC [B, IH_in, IW_in, lda] (channel slice base)
  int lda = 0;
  int M = 0, N = 0, K = 0;  // conv: M = B*OH*OW, K = 9*Cin
  int B = 0, OH = 0, OW = 0;  // conv output grid
  const void* W = nullptr;  // [N, K]
  const float* bias = nullptr;
  const float* gamma = nullptr;
  const void* resid = nullptr;
  int ld_resid = 0, resid_mod = 0;
  int act = ACT_NONE;
  void* out = nullptr;
  int ld_out = 0;
  int out_fp32 = 0;
  int rows_in = ROWS_PLAIN, remap_rows = 0, shuffle_cout = 0, IH = 0, IW = 0;
  int hm_S = 0, hm_heads = 0, hm_slices = 0;   // head-major store (see GemmArgs)
  const uint8_t* row_zero = nullptr;
  float2* stats_out = nullptr;          // producer: buffer with room for M * 2 * n_tiles float2
  const float2* stats_in = nullptr;     // consumer
  int stats_parts_in = 0;
  const float* colsum = nullptr;
  int ln_C = 0;
  float ln_eps = 0.f;
};

struct GemmOp {
  CUtensorMap ta, tb;
  GemmArgs args;
  int dtype, bn;
  unsigned grid;
  size_t smem;
  double flops;
};

// Returns 0 on success; on failure fills *err.
int gemm_build(const GemmDesc& d, GemmOp* op, std::string* err);
int gemm_launch(const GemmOp& op, cudaStream_t stream);

}  // namespace lwb
