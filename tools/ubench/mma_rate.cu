// Micro-benchmark: sustained tcgen05.mma rate (SS mode, K-major 128B-swizzled operands resident in shared memory),
// cta_group::1, one CTA per SM.  Prints cycles per MMA instruction and MAC/clk/SM for several N.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lw-detr_b200/csrc/ptx.cuh"
using namespace lwb;

template <int N>
__global__ void __launch_bounds__(128, 1) k_mma(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // halves 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc_k128(smem_u32(smem)), bd = umma_desc_k128(smem_u32(smem + 16384));
    const uint32_t idesc = umma_idesc_f16(false, 128, N);
    for (int w = 0; w < 8; ++w) umma_f16_ss(tm, ad, bd, idesc, 1);
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tm, ad + 2 * k, bd + 2 * k, idesc, 1);
    }
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

template <int N> void run(int iters) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_mma<N><<<148, 128, 1024 + 16384 + N * 128>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per = avg / (4.0 * iters);
  printf("cta_group::1 M=128 N=%3d K=16: %7.1f cycles/MMA  -> %7.0f MAC/clk/SM  (%s)\n", N, per, 128.0 * N * 16 / per, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<64>(2000); run<128>(2000); run<192>(2000); run<256>(2000);
  return 0;
}
