// Micro-benchmark (to run once at the start of a round, one gpurun call): the on-chip rates the attention kernel and
// the GEMM epilogue are designed around.
//   1. tcgen05.ld throughput per SM for the 32x32b shape at x16 / x32 / x64 / x128 with 4, 8 and 16 warps
//      (DESIGN.md 3.1 infers 64 B/clk/SM from kernel timings - this measures it directly)
//   2. tcgen05.st throughput, same sweep (x16 / x32)
//   3. cycles per tcgen05.mma (cta_group::1, M = 128, K = 16) for N = 16 ... 256 with the A operand in shared
//      memory and in TMEM (the P V product of attention is 8 such MMAs with N = head dim per 128 keys)
//   4. mbarrier hand-off: arrive in one warp -> try_wait returns in another -> arrive back (round-trip cycles),
//      and tcgen05.commit -> waiting warp released
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/tmem_rate tools/ubench/tmem_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lw-detr_b200/csrc/ptx.cuh"
using namespace lwb;

// ------------------------------------------------------------------------------------------------ 1 + 2: TMEM ld / st
template <int X> __device__ __forceinline__ void ld_x(uint32_t a, uint32_t& sink);
template <> __device__ __forceinline__ void ld_x<16>(uint32_t a, uint32_t& sink) {
  float v[16];
  tmem_ld_x16(a, v);
  tmem_ld_wait();
  sink ^= __float_as_uint(v[0]) ^ __float_as_uint(v[15]);
}
template <> __device__ __forceinline__ void ld_x<32>(uint32_t a, uint32_t& sink) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(a)
      : "memory");
  tmem_ld_wait();
  sink ^= r[0] ^ r[31];
}
// x64 / x128 as 2 / 4 back-to-back x32 loads with ONE wait (how the kernels issue them)
template <> __device__ __forceinline__ void ld_x<64>(uint32_t a, uint32_t& sink) {
  uint32_t r0, r1;
  asm volatile(
      "{\n\t.reg .b32 t<64>;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,"
      "t25,t26,t27,t28,t29,t30,t31}, [%2];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {t32,t33,t34,t35,t36,t37,t38,t39,t40,t41,t42,t43,t44,t45,t46,t47,t48,t49,t50,t51,t52,t53,"
      "t54,t55,t56,t57,t58,t59,t60,t61,t62,t63}, [%3];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      "mov.b32 %0, t0;\n\tmov.b32 %1, t63;\n\t}"
      : "=r"(r0), "=r"(r1)
      : "r"(a), "r"(a + 32)
      : "memory");
  sink ^= r0 ^ r1;
}
template <> __device__ __forceinline__ void ld_x<128>(uint32_t a, uint32_t& sink) {
  ld_x<64>(a, sink);          // two waits: upper bound on the cost of a 128-column pull
  ld_x<64>(a + 64, sink);
}

template <int X>
__global__ void __launch_bounds__(512, 1) k_ld(long long* out, int iters, int warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int i = 0; i < iters; ++i) ld_x<X>(tm + ((i * X) & 255), sink);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) out[148] = sink;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int X>
__global__ void __launch_bounds__(512, 1) k_st(long long* out, int iters, int warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < X / 16; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(tm + ((i * X + c * 16) & 255)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
                     "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                     : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

static double avg148(long long* d) {
  long long h[148];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  double a = 0;
  for (int i = 0; i < 148; ++i) a += h[i];
  return a / 148;
}

template <int X> void run_ld(long long* d) {
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    k_ld<X><<<148, 512>>>(d, iters, warps);
    cudaError_t e = cudaDeviceSynchronize();
    const double cyc = avg148(d);
    printf("tcgen05.ld 32x32b.x%-3d %2d warps: %8.1f cycles / (warp-load)   %7.1f B/clk/SM  (%s)\n", X, warps, cyc / iters,
           static_cast<double>(warps) * iters * 32 * X * 4 / cyc, cudaGetErrorString(e));
  }
}
template <int X> void run_st(long long* d) {
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    k_st<X><<<148, 512>>>(d, iters, warps);
    cudaError_t e = cudaDeviceSynchronize();
    const double cyc = avg148(d);
    printf("tcgen05.st 32x32b x%-3d  %2d warps: %8.1f cycles / (warp-store)  %7.1f B/clk/SM  (%s)\n", X, warps, cyc / iters,
           static_cast<double>(warps) * iters * 32 * X * 4 / cyc, cudaGetErrorString(e));
  }
}

// ------------------------------------------------------------------------------------------------ 3: small-N MMA cost
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc)
      : "memory");
}
template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k_mma(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc_k128(smem_u32(smem)), bd = umma_desc_k128(smem_u32(smem + 16384));
    const uint32_t idesc = umma_idesc_f16(false, 128, N);
    for (int w = 0; w < 8; ++w) umma_f16_ss(tm, ad, bd, idesc, 1);
    umma_commit(&bar); mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {                       // 8 K-steps of 16: the PV product of one 128-key chunk
        if (A_TMEM) mma_ts(tm, tm + 256 + 8 * k, bd + 2 * (k & 3), idesc);
        else umma_f16_ss(tm, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
      }
    }
    umma_commit(&bar); mbar_wait(&bar, 1);
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}
template <int N, bool A_TMEM> void run_mma(long long* d) {
  const int iters = 500;
  cudaFuncSetAttribute(k_mma<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_mma<N, A_TMEM><<<148, 128, 1024 + 16384 + 256 * 128>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  const double per = avg148(d) / (8.0 * iters);
  printf("tcgen05.mma M=128 N=%3d K=16 A in %s: %7.1f cycles/MMA  %6.0f MAC/clk/SM  (%s)\n", N, A_TMEM ? "TMEM" : "smem", per,
         128.0 * N * 16 / per, cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------ 4: hand-off latency
__global__ void __launch_bounds__(64, 1) k_handoff(long long* out, int iters) {
  __shared__ uint64_t ping, pong;
  if (threadIdx.x == 0) { mbar_init(&ping, 1); mbar_init(&pong, 1); fence_mbar_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(&ping);
      mbar_wait(&pong, i & 1);
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  } else {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&ping, i & 1);
      if (lane == 0) mbar_arrive(&pong);
    }
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 256 * 8);
  run_ld<16>(d); run_ld<32>(d); run_ld<64>(d); run_ld<128>(d);
  run_st<16>(d); run_st<32>(d);
  run_mma<16, false>(d); run_mma<32, false>(d); run_mma<64, false>(d); run_mma<128, false>(d); run_mma<256, false>(d);
  run_mma<16, true>(d); run_mma<32, true>(d); run_mma<64, true>(d); run_mma<128, true>(d);
  {
    const int iters = 2000;
    k_handoff<<<148, 64>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mbarrier ping-pong between two warps: %7.1f cycles per round trip  (%s)\n", avg148(d) / iters, cudaGetErrorString(e));
  }
  cudaFree(d);
  return 0;
}
