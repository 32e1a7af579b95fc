// tcgen05 / TMA attention for long sequences (the ViT global-attention blocks, vit.py:201-204 + 130-137).
//
// One CTA owns 128 query rows of one (sequence, head).  Q, K and V tiles are fetched by TMA straight out of the
// packed [rows, 3C] qkv matrix (box = 128 rows x dh columns at column  part*C + head*dh, hardware swizzle
// matching the dh*2-byte row pitch), S = Q K^T and O = P V run on the tensor core with accumulators in TMEM:
//
//   warp 0   : TMA producer (Q once, then a ring of K / V key chunks of 128 keys)
//   warp 1   : TMEM allocator + tcgen05.mma issuer.  S: A = Q (smem, K-major), B = K chunk (smem, K-major);
//              O: A = P (TMEM, written by the softmax warps), B = V chunk (smem, MN-major)
//   warps 2-5: softmax, one thread per query row (no shuffles): tcgen05.ld S row -> exp2 -> P (16-bit) via
//              tcgen05.st, row sum in registers; finally O / l -> global.
//
// Two passes over the keys instead of an online rescale: pass A only takes the row maxima of S, pass B
// recomputes S, exponentiates against the final maximum and accumulates O - so O never needs a correction
// step in TMEM.  QK^T is cheap at these head dims (1-4 K-steps); the kernel is bound by exp (MUFU) at
// dh <= 32 and close to it at dh = 64.  TMEM budget: S 128 + P 64 + O dh <= 256 columns -> two CTAs per SM.
#include "attn.h"
#include "ptx.cuh"
#include "tma_util.h"

#include <algorithm>

namespace lwb {

static constexpr int TC_BM = 128;     // query rows per CTA
static constexpr int TC_BKV = 128;    // keys per chunk
static constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, int DH>
__global__ void __launch_bounds__(TC_THREADS) attn_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttnArgs p, int C) {
  constexpr int TC_NS = DH == 64 ? 2 : 3;                      // K and V ring depth (two CTAs per SM must fit)
  constexpr int TILE_BYTES = 128 * DH * 2;                     // one 128-row tile of Q, K or V
  constexpr uint32_t PITCH = DH * 2;                           // bytes per row = swizzle span
  constexpr uint32_t LAYOUT = DH == 64 ? 2u : (DH == 32 ? 4u : 6u);   // SWIZZLE_128B / 64B / 32B
  constexpr uint32_t SBO = 8 * PITCH;                          // 8-row group stride (K-major and MN-major alike)
  constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;      // TMEM column map (256 columns allocated)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TC_NS * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + TC_NS * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                 // [NS]
  uint64_t* k_empty = k_full + TC_NS;          // [NS]
  uint64_t* v_full = k_empty + TC_NS;          // [NS]
  uint64_t* v_empty = v_full + TC_NS;          // [NS]
  uint64_t* s_full = v_empty + TC_NS;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int q0 = qtile * TC_BM;
  const int row0 = seq * p.seqlen;                              // first matrix row of this sequence
  const int nchunks = (p.seqlen + TC_BKV - 1) / TC_BKV;
  const int colq = head * DH, colk = C + head * DH, colv = 2 * C + head * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < TC_NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(sQ, &tm, q_full, colq, row0 + q0);
      uint32_t kc = 0, vc = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nchunks; ++j) {
          {
            const int s = kc % TC_NS;
            mbar_wait(&k_empty[s], ((kc / TC_NS) & 1) ^ 1);
            mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
            tma_load_2d(sK + s * TILE_BYTES, &tm, &k_full[s], colk, row0 + j * TC_BKV);
            ++kc;
          }
          if (pass == 1) {
            const int s = vc % TC_NS;
            mbar_wait(&v_empty[s], ((vc / TC_NS) & 1) ^ 1);
            mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
            tma_load_2d(sV + s * TILE_BYTES, &tm, &v_full[s], colv, row0 + j * TC_BKV);
            ++vc;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr bool BF = Cvt<T>::is_bf16;
      constexpr uint32_t idesc_s = umma_idesc_f16(BF, 128, TC_BKV);                    // S: N = 128 keys, both K-major
      constexpr uint32_t idesc_o = umma_idesc_f16(BF, 128, DH) | (1u << 16);           // O: B (= V) is MN-major
      uint32_t kc = 0, vc = 0, sc = 0, pc = 0;
      const uint64_t qdesc = umma_desc(smem_u32(sQ), SBO, LAYOUT);
      mbar_wait(q_full, 0);
      auto issue_s = [&]() {
        const int s = kc % TC_NS;
        mbar_wait(&k_full[s], (kc / TC_NS) & 1);
        mbar_wait(s_empty, (sc & 1) ^ 1);           // softmax has finished reading the previous S tile
        tc_fence_after();
        const uint64_t kdesc = umma_desc(smem_u32(sK + s * TILE_BYTES), SBO, LAYOUT);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) umma_f16_ss(tmem + COL_S, qdesc + 2 * kk, kdesc + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
        ++kc;
        ++sc;
      };
      for (int j = 0; j < nchunks; ++j) issue_s();                 // pass A: maxima only
      issue_s();                                                   // pass B, chunk 0
      for (int j = 0; j < nchunks; ++j) {
        if (j + 1 < nchunks) issue_s();                            // S_{j+1} overlaps the softmax of chunk j+1's wait
        const int s = vc % TC_NS;
        mbar_wait(p_full, pc & 1);
        mbar_wait(&v_full[s], (vc / TC_NS) & 1);
        tc_fence_after();
        const uint64_t vdesc = umma_desc(smem_u32(sV + s * TILE_BYTES), SBO, LAYOUT);
#pragma unroll
        for (int kk = 0; kk < TC_BKV / 16; ++kk)                   // 16 keys per MMA: A advances 8 TMEM columns, B 16 rows
          umma_f16_ts(tmem + COL_O, tmem + COL_P + 8 * kk, vdesc + ((16 * PITCH) >> 4) * kk, idesc_o, (j | kk) != 0 ? 1u : 0u);
        umma_commit(&v_empty[s]);
        umma_commit(p_empty);
        ++vc;
        ++pc;
      }
      umma_commit(o_full);
    }
  } else {
    // -------------------------------------------------------------------- softmax / epilogue: thread = query row
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t sc = 0, pc = 0;
    float m = -INFINITY;
    for (int j = 0; j < nchunks; ++j) {                            // ---- pass A: row maxima
      mbar_wait(s_full, sc & 1);
      tc_fence_after();
      const int kvalid = p.seqlen - j * TC_BKV;                    // keys >= kvalid belong to the next sequence
#pragma unroll 1
      for (int piece = 0; piece < 4; ++piece) {
        float v[32];
        __syncwarp();
        tmem_ld_x32(trow + COL_S + piece * 32, v);
        tmem_ld_wait();
        if (kvalid >= (piece + 1) * 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (piece * 32 + i < kvalid) m = fmaxf(m, v[i]);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      ++sc;
    }
    const float msc = m * p.scale_log2;
    float l = 0.f;
    for (int j = 0; j < nchunks; ++j) {                            // ---- pass B: P = exp2(S*scale - m*scale), row sums
      mbar_wait(s_full, sc & 1);
      mbar_wait(p_empty, (pc & 1) ^ 1);                            // the previous P tile has been consumed by its MMA
      tc_fence_after();
      const int kvalid = p.seqlen - j * TC_BKV;
#pragma unroll 1
      for (int piece = 0; piece < 4; ++piece) {
        float v[32];
        __syncwarp();
        tmem_ld_x32(trow + COL_S + piece * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        if (kvalid >= (piece + 1) * 32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = ex2f(fmaf(v[2 * i], p.scale_log2, -msc));
            const float b = ex2f(fmaf(v[2 * i + 1], p.scale_log2, -msc));
            l += a + b;
            pk[i] = Cvt<T>::pack(a, b);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = (piece * 32 + 2 * i < kvalid) ? ex2f(fmaf(v[2 * i], p.scale_log2, -msc)) : 0.f;
            const float b = (piece * 32 + 2 * i + 1 < kvalid) ? ex2f(fmaf(v[2 * i + 1], p.scale_log2, -msc)) : 0.f;
            l += a + b;
            pk[i] = Cvt<T>::pack(a, b);
          }
        }
        __syncwarp();
        tmem_st_x16(trow + COL_P + piece * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(s_empty);
      mbar_arrive(p_full);
      ++sc;
      ++pc;
    }
    // ---- O / l -> global
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int qrow = q0 + r;
    T* dst = reinterpret_cast<T*>(p.o) + (static_cast<long long>(row0) + qrow) * p.ldo + head * DH;
#pragma unroll
    for (int c = 0; c < DH / 16; ++c) {
      float v[16];
      __syncwarp();
      tmem_ld_x16(trow + COL_O + c * 16, v);
      tmem_ld_wait();
      if (qrow < p.seqlen) {
        U8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = Cvt<T>::pack(v[2 * i] * inv, v[2 * i + 1] * inv);
        stg256(dst + c * 16, o);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <typename T, int DH>
static int launch_tc(const AttnArgs& a, int C, cudaStream_t st) {
  CUtensorMap tm;
  std::string err;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.ldq), static_cast<cuuint64_t>(a.nseq) * a.seqlen};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.ldq) * 2};
  const cuuint32_t box[2] = {DH, 128};
  if (tma_encode(&tm, Cvt<T>::is_bf16 ? DT_BF16 : DT_F16, 2, a.q, dims, strides, box, DH * 2, &err)) return -3;
  constexpr int TC_NS = DH == 64 ? 2 : 3;
  const size_t smem = 1024 + static_cast<size_t>(1 + 2 * TC_NS) * 128 * DH * 2 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  dim3 grid((a.seqlen + TC_BM - 1) / TC_BM, a.heads, a.nseq);
  attn_tc_kernel<T, DH><<<grid, TC_THREADS, smem, st>>>(tm, a, C);
  return static_cast<int>(cudaGetLastError());
}

// Packed-qkv fast path: q, k, v are the column blocks [0,C), [C,2C), [2C,3C) of one 16-bit matrix.
int attention_tc_launch(int dtype, const AttnArgs& a, int dh, int C, cudaStream_t st) {
  if (dtype == DT_BF16) {
    if (dh == 16) return launch_tc<__nv_bfloat16, 16>(a, C, st);
    if (dh == 32) return launch_tc<__nv_bfloat16, 32>(a, C, st);
    if (dh == 64) return launch_tc<__nv_bfloat16, 64>(a, C, st);
  } else {
    if (dh == 16) return launch_tc<__half, 16>(a, C, st);
    if (dh == 32) return launch_tc<__half, 32>(a, C, st);
    if (dh == 64) return launch_tc<__half, 64>(a, C, st);
  }
  return -2;
}

}  // namespace lwb
