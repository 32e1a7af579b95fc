// Multi-scale deformable attention core, forward (reference: models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:33-84 bilinear, :237-299 kernel, launcher :923-954; host glue
// ms_deform_attn_cuda.cu:20-80; module math ops/modules/ms_deform_attn.py:118-131).
//
// B200 design: an HBM/L2 gather, no tensor cores.  The reference spends one thread per output
// channel, re-reading the sampling location and weight D times and issuing scalar loads.  Here one
// thread owns 8 channels of one (query, head): every bilinear corner is ONE 16-byte read-only load
// (two adjacent threads cover the 32-byte head slice = one DRAM sector), the location arithmetic,
// the softmax over the L*P logits of the head and the weighted accumulation stay in registers, and
// the [B*nq, d] result is written with 16-byte coalesced stores.  The kernel consumes the raw
// sampling_offsets / attention_weights projections directly (softmax and location math fused), so
// the [B,nq,M,L,P,2] location tensor of the reference never exists in memory.
#include "msda.h"
#include "launch.h"
#include "ptx.cuh"

namespace lwb {

__device__ __forceinline__ U4 ldg_nc16(const void* p) {
  U4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <typename T, int NL, int NP>   // levels, points per head and level
__global__ void __launch_bounds__(256, (NL * NP <= 2) ? 8 : ((NL * NP <= 4) ? 6 : 4)) msda_fwd_kernel(const MsdaArgs p) {
  pdl_sync();   // programmatic dependent launch: release the successor, wait for the predecessor (launch.h)
  constexpr int D = 16;
  constexpr int LP = NL * NP;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int half = static_cast<int>(gid & 1);
  const long long hm = gid >> 1;
  const int m = static_cast<int>(hm % p.heads);
  const long long row = hm / p.heads;                 // b * nq + q
  if (row >= static_cast<long long>(p.batch) * p.nq) return;
  const int b = static_cast<int>(row / p.nq);

  // ---- per-(query, head) scalars: 2*LP offsets, LP logits (16-bit), reference box (fp32)
  const T* oa = reinterpret_cast<const T*>(p.offs_logits) + row * p.ld_ol;
  float off[2 * LP], w[LP];
  {
    const uint32_t* o32 = reinterpret_cast<const uint32_t*>(oa + m * (2 * LP));
#pragma unroll
    for (int i = 0; i < LP; ++i) {
      const float2 f = Cvt<T>::unpack(__ldg(o32 + i));
      off[2 * i] = f.x;
      off[2 * i + 1] = f.y;
    }
    const T* lg = oa + p.heads * (2 * LP) + m * LP;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < LP; ++i) {
      w[i] = Cvt<T>::to_f(lg[i]);
      mx = fmaxf(mx, w[i]);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LP; ++i) {
      w[i] = __expf(w[i] - mx);
      sum += w[i];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < LP; ++i) w[i] *= inv;
  }
  const float4 ref = __ldg(reinterpret_cast<const float4*>(p.ref) + row);
  const float sx = ref.z * (0.5f / NP);
  const float sy = ref.w * (0.5f / NP);

  const T* vbase = reinterpret_cast<const T*>(p.value) + static_cast<long long>(b) * p.S * p.ldv + m * D + half * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

#pragma unroll
  for (int i = 0; i < LP; ++i) {
    const int l = i / NP;
    const int H = p.lvl_h[l], W = p.lvl_w[l];
    const float x = (ref.x + off[2 * i] * sx) * W - 0.5f;      // ms_deform_attn.py:125-127, cuh:285-286
    const float y = (ref.y + off[2 * i + 1] * sy) * H - 0.5f;
    if (!(y > -1.f && x > -1.f && y < H && x < W)) continue;   // cuh:288
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const float lx = x - xf, ly = y - yf;
    const T* vl = vbase + static_cast<long long>(p.lvl_start[l]) * p.ldv;
    const float wi = w[i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xi = x0 + (c & 1), yi = y0 + (c >> 1);
      if (xi < 0 || yi < 0 || xi >= W || yi >= H) continue;
      const float cw = wi * ((c & 1) ? lx : 1.f - lx) * ((c >> 1) ? ly : 1.f - ly);
      const U4 v = ldg_nc16(vl + static_cast<long long>(yi * W + xi) * p.ldv);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = Cvt<T>::unpack(u[j]);
        acc[2 * j] = fmaf(cw, f.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(cw, f.y, acc[2 * j + 1]);
      }
    }
  }
  U4 o;
  o.x = Cvt<T>::pack(acc[0], acc[1]);
  o.y = Cvt<T>::pack(acc[2], acc[3]);
  o.z = Cvt<T>::pack(acc[4], acc[5]);
  o.w = Cvt<T>::pack(acc[6], acc[7]);
  *reinterpret_cast<U4*>(reinterpret_cast<T*>(p.out) + row * p.ld_out + m * D + half * 8) = o;
}

template <typename T>
static int dispatch(const MsdaArgs& a, cudaStream_t st) {
  const long long threads = static_cast<long long>(a.batch) * a.nq * a.heads * 2;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  if (a.levels == 1 && a.points == 2) launch_k(msda_fwd_kernel<T, 1, 2>, dim3(grid), dim3(256), 0, st, a);
  else if (a.levels == 2 && a.points == 4) launch_k(msda_fwd_kernel<T, 2, 4>, dim3(grid), dim3(256), 0, st, a);
  else if (a.levels == 1 && a.points == 4) launch_k(msda_fwd_kernel<T, 1, 4>, dim3(grid), dim3(256), 0, st, a);
  else if (a.levels == 2 && a.points == 2) launch_k(msda_fwd_kernel<T, 2, 2>, dim3(grid), dim3(256), 0, st, a);
  else if (a.levels == 4 && a.points == 4) launch_k(msda_fwd_kernel<T, 4, 4>, dim3(grid), dim3(256), 0, st, a);
  else return -2;
  return static_cast<int>(cudaGetLastError());
}

int msda_launch(int dtype, const MsdaArgs& a, cudaStream_t st) {
  return dtype == DT_BF16 ? dispatch<__nv_bfloat16>(a, st) : dispatch<__half>(a, st);
}

}  // namespace lwb
