// CTA-level GEMM engine on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// with an error-compensated BF16x3 split so that the result meets the 1e-4 per-step parity bar:
//
//     a = a_hi + a_lo  (both bf16, round-to-nearest),   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
//
// (relative error ~1e-5, SURVEY.md section 7 "hard part 1"; single-pass BF16/TF32 is 3e-3 / 4e-4.)
//
//   C[m][n] = sum_k A(m,k) * B(n,k)        same contract and operand addressing as gemm() (nmb_gemm.cuh)
//
// Data flow per 64-wide K chunk: all 256 threads read the FP32 operands from global memory (L2),
// split them into hi/lo BF16 planes and write them to shared memory directly in the UMMA canonical
// *no-swizzle* layouts -- K-major for k-contiguous sources, MN-major for i-contiguous sources, so no
// transposition is ever needed -- then one thread issues the tcgen05.mma instructions (3 per K=16
// step and 128-row accumulator) and commits them to an mbarrier.  Up to two 128 x BN accumulators
// (256 batch rows) live in TMEM and share the staged B operand.  The epilogue reads TMEM with
// tcgen05.ld (one accumulator row per thread, 16 columns at a time) and applies the stage's functor.
//
// Two CTAs are resident per SM (2 x 96.5 KB shared memory, 2 x 256 TMEM columns): while one CTA
// stages operands on the CUDA cores the other one's MMAs occupy the tensor pipe.
#pragma once
#include <cuda_bf16.h>

#include "nmb_gemm.cuh"

namespace nmb {
namespace tc {

constexpr int TBK = 64;                       // K elements per staged chunk
constexpr int A_ROWS = 256;                   // staged A rows = two 128-row accumulators
constexpr int B_ROWS = 128;                   // max N per tile (TMEM columns per accumulator)
constexpr int A_LBO = (A_ROWS + 1) * 16;      // bytes between consecutive 8-wide k groups (+16: bank spread)
constexpr int B_LBO = (B_ROWS + 1) * 16;
constexpr int SBO = 128;                      // bytes between consecutive 8-row (K-major) / 8-column (MN-major) groups
constexpr int A_PLANE = (TBK / 8) * A_LBO;    // 32 896 B
constexpr int B_PLANE = (TBK / 8) * B_LBO;    // 16 512 B
constexpr int kSmemBytes = 2 * A_PLANE + 2 * B_PLANE;   // hi + lo planes of A and B = 98 816 B
constexpr int kTmemCols = 256;

struct Ctx {
  unsigned char* smem;     // kSmemBytes, 16-byte aligned
  uint64_t* mbar;          // one mbarrier (shared memory)
  uint32_t tmem_base;      // from tcgen05.alloc
  uint32_t phase;          // parity of the next mbarrier completion
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done) : "r"(addr), "r"(phase) : "memory");
    if (!done && clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {     // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols) {           // one full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

// Instruction descriptor, kind::f16: D = F32, A = B = BF16, M = 128 (cute::UMMA::InstrDescriptor).
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Accumulator load split into issue and wait, so that independent work overlaps the TMEM read.  The wait names
// the destination registers as in/out operands: no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}

// Two FP32 values -> packed bf16 "hi" pair and bf16 "lo" pair (element a in the low half): 6 instructions
// (2 F2FP, shift, mask, 2 FADD); bit-identical to __floats2bfloat162_rn on (a, b) and on the residuals.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}

// 8 FP32 values -> 8 bf16 "hi" + 8 bf16 "lo" (16 bytes each).
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
}

// 8 consecutive floats at src[0..8) of which only the first `valid` exist (16-byte aligned src).
__device__ __forceinline__ void load8(const float* src, int valid, float (&x)[8]) {
  if (valid >= 8) {
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = i < valid ? src[i] : 0.f;
  }
}

// Stage `rows` x `kc` elements of operand o (tile origin i0, k0) into the hi/lo planes.
//   k-major source  -> UMMA K-major  layout: byte(r,k) = (k/8)*LBO + r*16 + (k%8)*2
//   i-major source  -> UMMA MN-major layout: byte(i,k) = (k/8)*LBO + (i/8)*128 + (k%8)*16 + (i%8)*2
// Out-of-range rows / k are written as zeros (the MMA always reads whole 128 x 16 operand blocks).
__device__ __forceinline__ void stage(const Opnd& o, int i0, int I, int k0, int K, int rows, int kc,
                                      unsigned char* hi, unsigned char* lo, int lbo) {
  if (o.kmajor) {
    const int kg = kc >> 3;
    const int units = rows * kg;
    for (int u = threadIdx.x; u < units; u += kThreads) {
      const int g = u % kg, r = u / kg;
      const int gi = i0 + r, gk = k0 + (g << 3);
      float x[8];
      if (gi < I && gk < K) load8(o.p + (long long)gi * o.ld + gk, K - gk, x);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      uint4 h, l;
      split8(x, h, l);
      const int off = g * lbo + (r << 4);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  } else {
    const int rg = rows >> 3;
    const int units = kc * rg;
    for (int u = threadIdx.x; u < units; u += kThreads) {
      const int kk = u & 7, rest = u >> 3;
      const int h8 = rest % rg, kgi = rest / rg;
      const int gk = k0 + (kgi << 3) + kk, gi = i0 + (h8 << 3);
      float x[8];
      if (gk < K && gi < I) load8(o.p + (long long)gk * o.ld + gi, I - gi, x);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      uint4 h, l;
      split8(x, h, l);
      const int off = kgi * lbo + (h8 << 7) + (kk << 4);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

// All threads of the CTA must call with identical arguments.
// epi.rowc(m, n0, nvalid, v) receives C[m][n0 .. n0+16) (first nvalid entries are inside N).
template <class Epi>
__device__ __noinline__ void gemm(int M, int N, int K, Opnd A, Opnd B, Epi& epi, Ctx& c) {
  unsigned char* a_hi = c.smem;
  unsigned char* a_lo = a_hi + A_PLANE;
  unsigned char* b_hi = a_lo + A_PLANE;
  unsigned char* b_lo = b_hi + B_PLANE;
  const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int m0 = 0; m0 < M; m0 += A_ROWS) {
    const int n_mt = (M - m0 > 128) ? 2 : 1;
    for (int n0 = 0; n0 < N; n0 += B_ROWS) {
      const int bn = min(B_ROWS, (N - n0 + 15) & ~15);
      const uint32_t idesc = make_idesc(bn, !A.kmajor, !B.kmajor);
      for (int k0 = 0; k0 < K; k0 += TBK) {
        const int kc = min(TBK, (K - k0 + 15) & ~15);
        stage(A, m0, M, k0, K, n_mt * 128, kc, a_hi, a_lo, A_LBO);
        stage(B, n0, N, k0, K, bn, kc, b_hi, b_lo, B_LBO);
        fence_async_smem();               // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        if (threadIdx.x == 0) {
          fence_after();
          for (int ks = 0; ks < (kc >> 4); ++ks) {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {        // hi*hi, lo*hi, hi*lo
              const uint32_t ab = (pass == 1 ? sa_lo : sa_hi) + ks * 2 * A_LBO;
              const uint64_t bd = make_desc((pass == 2 ? sb_lo : sb_hi) + ks * 2 * B_LBO, B_LBO);
              const uint32_t acc = (k0 > 0 || ks > 0 || pass > 0) ? 1u : 0u;
              for (int mt = 0; mt < n_mt; ++mt)
                mma_bf16(c.tmem_base + mt * B_ROWS, make_desc(ab + mt * 128 * 16, A_LBO), bd, idesc, acc);
            }
          }
          mma_commit(c.mbar);             // arrives when every MMA above has finished reading smem / writing TMEM
        }
        mbar_wait(c.mbar, c.phase);
        c.phase ^= 1u;
      }
      fence_after();
      // epilogue: warp w owns TMEM lanes (w%4)*32 .. +31
      const int row_in_tile = ((warp & 3) << 5) + lane;
      const int ncg = (bn + 15) >> 4;
      if (n_mt == 2) {
        const int mt = warp >> 2;
        const int m = m0 + mt * 128 + row_in_tile;
        for (int cg = 0; cg < ncg; ++cg) {
          float v[16];
          tmem_ld16(c.tmem_base + ((uint32_t)((warp & 3) << 5) << 16) + mt * B_ROWS + (cg << 4), v);
          const int nn = n0 + (cg << 4);
          if (m < M && nn < N) epi.rowc(m, nn, min(16, N - nn), v);
        }
      } else {
        const int m = m0 + row_in_tile;
        for (int cg = warp >> 2; cg < ncg; cg += 2) {   // warps w and w+4 share rows, alternate column groups
          float v[16];
          tmem_ld16(c.tmem_base + ((uint32_t)((warp & 3) << 5) << 16) + (cg << 4), v);
          const int nn = n0 + (cg << 4);
          if (m < M && nn < N) epi.rowc(m, nn, min(16, N - nn), v);
        }
      }
      fence_before();
      __syncthreads();                    // TMEM and smem are free for the next tile / stage
    }
  }
}

}  // namespace tc
}  // namespace nmb
