// Micro-benchmark: TMA -> mbarrier -> tcgen05.mma -> commit -> mbarrier ring (no epilogue), one CTA per SM.
// Each CTA streams `nkb` k-blocks: A box 128 rows x 64 cols (and optionally a W box N x 64) per k-block.
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
#include "../../lw-detr_b200/csrc/ptx.cuh"
#include "../../lw-detr_b200/csrc/tma_util.h"
#include "../../lw-detr_b200/csrc/gemm_tc.h"
using namespace lwb;

template <int N>
__global__ void __launch_bounds__(128, 1) k_pipe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                 long long* out, int nkb, int stages, int load_b, int rows_total, int kcols) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem; uint8_t* sB = smem + stages * 16384;
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&done, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages; const uint32_t ph = (i / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], 16384 + (load_b ? N * 128 : 0));
      const int kblocks = kcols / 64;
      const int row = ((blockIdx.x + (i / kblocks) * gridDim.x) * 128) % rows_total;
      tma_load_2d(sA + s * 16384, &tmA, &full[s], (i % kblocks) * 64, row);
      if (load_b) tma_load_2d(sB + s * N * 128, &tmB, &full[s], (i % kblocks) * 64, 0);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(false, 128, N);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages; const uint32_t ph = (i / stages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint64_t ad = umma_desc_k128(smem_u32(sA + s * 16384)), bd = umma_desc_k128(smem_u32(sB + (load_b ? s : 0) * N * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tm, ad + 2 * k, bd + 2 * k, idesc, 1);
      umma_commit(&empty[s]);
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

int main() {
  const int M = 51200, N = 192;
  for (int kcols : {192, 768}) {
    void *A, *W; cudaMalloc(&A, (size_t)M * kcols * 2); cudaMalloc(&W, (size_t)N * kcols * 2);
    cudaMemset(A, 0, (size_t)M * kcols * 2); cudaMemset(W, 0, (size_t)N * kcols * 2);
    CUtensorMap ta, tb; std::string err;
    cuuint64_t da[2] = {(cuuint64_t)kcols, (cuuint64_t)M}, sa[1] = {(cuuint64_t)kcols * 2}; cuuint32_t ba[2] = {64, 128};
    cuuint64_t db[2] = {(cuuint64_t)kcols, (cuuint64_t)N}; cuuint32_t bb[2] = {64, (cuuint32_t)N};
    if (tma_encode(&ta, DT_F16, 2, A, da, sa, ba, 128, &err) || tma_encode(&tb, DT_F16, 2, W, db, sa, bb, 128, &err)) { printf("encode: %s\n", err.c_str()); return 1; }
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k_pipe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int load_b : {0, 1}) for (int stages : {2, 3, 4, 5}) {
      const int nkb = 240;
      const size_t smem = 1024 + (size_t)stages * (16384 + N * 128);
      for (int rep = 0; rep < 2; ++rep) k_pipe<N><<<148, 128, smem>>>(ta, tb, d, nkb, stages, load_b, M, kcols);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("K=%4d load_W=%d stages=%d : %7.0f cycles per k-block (MMA floor 384)  [%s]\n", kcols, load_b, stages, avg / nkb, cudaGetErrorString(e));
    }
    cudaFree(A); cudaFree(W); cudaFree(d);
  }
  return 0;
}
