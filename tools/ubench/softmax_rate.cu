// Micro-benchmark: instruction-level cost of the softmax inner loop of the attention kernels, on registers only
// (no TMEM, no tensor core): how many clocks per score per SM do the candidate formulations need, and how far
// below the MUFU floor (16 ex2/clk/SM = 0.0625 clk/score) can a polynomial exp2 on the FMA pipe take them?
//   MODE 0: scalar FFMA + ex2.approx + pack + FADD row sums          (round-1 inner loop)
//   MODE 1: packed FFMA2 (fma.rn.f32x2) + ex2.approx + pack, row sums left to the PV product
//   MODE 2: MODE 1 with the pairs selected by PMASK (bit i = pair i of every 8 pairs) on a degree-3 polynomial
//           exp2 built from FFMA2 / FADD2 (Cody-Waite split by the 1.5*2^23 magic add) instead of the MUFU
// Every variant also runs the 3-input row maximum over its 64 scores and stores the packed P (st.shared.v4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/softmax_rate tools/ubench/softmax_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

static constexpr float MAGIC = 12582912.f;   // 1.5 * 2^23

// exp2 of the pair (s.x*c - m, s.y*c - m) without the MUFU
// (the scores are clamped from below first: an argument under -126 would leave the fraction outside the polynomial's range)
__device__ __forceinline__ void exp2_poly_pair(uint64_t s2in, uint64_t c2, uint64_t magic_minus_m2, uint64_t negm2, float& e0, float& e1) {
  float s0, s1;
  upk2(s2in, s0, s1);
  const uint64_t s2 = pk2(fmaxf(s0, -300.f), fmaxf(s1, -300.f));
  const uint64_t t2 = fma2(s2, c2, magic_minus_m2);          // round(s*c - m) + MAGIC
  float t0, t1;
  upk2(t2, t0, t1);
  const uint64_t r2 = add2(t2, pk2(-MAGIC, -MAGIC));         // the integer part
  const uint64_t u2 = fma2(r2, pk2(-1.f, -1.f), negm2);      // -m - r
  const uint64_t f2 = fma2(s2, c2, u2);                      // fraction in [-0.5, 0.5]
  uint64_t p2 = fma2(f2, pk2(0.05517164617776871f, 0.05517164617776871f), pk2(0.2426111251115799f, 0.2426111251115799f));
  p2 = fma2(p2, f2, pk2(0.6932609677314758f, 0.6932609677314758f));
  p2 = fma2(p2, f2, pk2(0.9999280571937561f, 0.9999280571937561f));
  float p0, p1;
  upk2(p2, p0, p1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int MODE, uint32_t PMASK>
__global__ void __launch_bounds__(512, 1) k_softmax(const float* in, float* out, long long* clk, int iters, float c) {
  extern __shared__ uint4 sm[];
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = in[(threadIdx.x * 64 + i) & 4095];
  float lsum[4] = {0.f, 0.f, 0.f, 0.f};
  float mrun = -1e30f;
  uint4* dst = sm + threadIdx.x * 8;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // row maximum of the chunk (3-input min/max), nothing loop-invariant: the reference moves every iteration
    float mm[4] = {mrun, mrun, mrun, mrun};
#pragma unroll
    for (int i = 0; i < 16; i += 2)
#pragma unroll
      for (int q = 0; q < 4; ++q) mm[q] = fmaxf(mm[q], fmaxf(v[q * 16 + i], v[q * 16 + i + 1]));
    const float mx = fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])) + 1e-3f * it;
    mrun = mx * 0.5f;
    const float msc = mx * c;
    uint32_t pk[32];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e0 = ex2f(fmaf(v[2 * i], c, -msc)), e1 = ex2f(fmaf(v[2 * i + 1], c, -msc));
        pk[i] = pack_h2(e0, e1);
        lsum[i & 3] += e0 + e1;
      }
    } else {
      const uint64_t c2 = pk2(c, c), nm2 = pk2(-msc, -msc), mg2 = pk2(MAGIC - msc, MAGIC - msc);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint64_t s2 = pk2(v[2 * i], v[2 * i + 1]);
        float e0, e1;
        if (MODE == 2 && ((PMASK >> (i & 7)) & 1u)) {
          exp2_poly_pair(s2, c2, mg2, nm2, e0, e1);
        } else {
          float a0, a1;
          upk2(fma2(s2, c2, nm2), a0, a1);
          e0 = ex2f(a0);
          e1 = ex2f(a1);
        }
        pk[i] = pack_h2(e0, e1);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    // feed a data dependence back so that iterations cannot be merged
    v[0] += __uint_as_float(pk[5] & 0x3f800000u) * 1e-6f;
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = lsum[0] + lsum[1] + lsum[2] + lsum[3] + v[5] + mrun + __uint_as_float(dst[3].x);
}

// packed-vs-scalar FMA issue rate
template <int PACKED>
__global__ void __launch_bounds__(512, 1) k_fma(float* out, long long* clk, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint64_t x = pk2(a[2 * i], a[2 * i + 1]);
        x = fma2(x, pk2(0.999f, 0.999f), pk2(1e-3f, 1e-3f));
        upk2(x, a[2 * i], a[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 0.999f, 1e-3f);
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, uint32_t PMASK>
static void run(const char* name, const float* in, float* out, long long* clk, int warps) {
  const int iters = 2000;
  cudaFuncSetAttribute(k_softmax<MODE, PMASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 128);
  for (int rep = 0; rep < 2; ++rep) k_softmax<MODE, PMASK><<<148, warps * 32, 512 * 128>>>(in, out, clk, iters, 0.36f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per_score = double(mx) / iters / (double(warps) * 32 * 64);
  printf("%-44s %2d warps: %8.1f clk/iter  %.4f clk/score/SM  (MUFU floor 0.0625; x%.2f)  [%s]\n", name, warps, double(mx) / iters, per_score,
         0.0625 / per_score, cudaGetErrorString(e));
}

int main() {
  float *in, *out;
  long long* clk;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&clk, 148 * 16 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -8.f + 16.f * ((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  for (int warps : {8, 16}) {
    run<0, 0>("scalar FFMA + MUFU + FADD sums", in, out, clk, warps);
    run<1, 0>("FFMA2 + MUFU (sums via MMA)", in, out, clk, warps);
    run<2, 0x11>("FFMA2, 2/8 pairs polynomial", in, out, clk, warps);
    run<2, 0x49>("FFMA2, 3/8 pairs polynomial", in, out, clk, warps);
    run<2, 0x55>("FFMA2, 4/8 pairs polynomial", in, out, clk, warps);
    run<2, 0x6d>("FFMA2, 5/8 pairs polynomial", in, out, clk, warps);
    run<2, 0xff>("FFMA2, all polynomial", in, out, clk, warps);
  }
  for (int packed : {0, 1}) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
      if (packed) k_fma<1><<<148, 512>>>(out, clk, iters); else k_fma<0><<<148, 512>>>(out, clk, iters);
    }
    cudaDeviceSynchronize();
    long long hc[16];
    cudaMemcpy(hc, clk, sizeof hc, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 16; ++i) mx = hc[i] > mx ? hc[i] : mx;
    printf("%s: %.2f FMA/clk/SM (16 warps, 16 independent chains)\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar", 16.0 * 512 * iters / double(mx));
  }
  return 0;
}
