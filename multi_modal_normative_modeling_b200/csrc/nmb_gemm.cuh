// CTA-level FP32 GEMM engine used by every dense stage of the fused cVAE kernels.
//
//   C[m][n] = sum_k A(m,k) * B(n,k)          m < M, n < N, k < K
//
// One CTA (256 threads) computes the whole product, tile by tile (128 x 16*TN), with both
// operands streamed from global memory (L2-resident) through double-buffered shared-memory
// tiles and an 8 x TN register tile per thread.  Each accumulator is handed to an epilogue
// functor, which is where bias/activation, the loss, leaky-relu', and the Adam update are fused.
//
// Operands are addressed as either "k-major"  elem(i,k) = p[i*ld + k]   (activations as A,
// weights as B in the forward pass) or "i-major" elem(i,k) = p[k*ld + i] (weights as B in the
// data-gradient, activations / output-gradients in the weight-gradient), so no transposed
// copies are ever materialised.  All leading dimensions are multiples of 4 floats and all
// base pointers 16-byte aligned, so every global access is a float4.
//
// FP32 FFMA is the parity-safe arithmetic: SURVEY.md measured single-pass TF32/BF16 tensor
// math at 3.7e-4 .. 6e-2 relative error on this network, outside the 1e-4 per-step bar.
#pragma once
#include "nmb_common.cuh"

namespace nmb {

constexpr int BM = 128;   // rows of C per tile
constexpr int BK = 16;    // k-slice per pipeline stage
constexpr int TM = 8;     // rows per thread
constexpr int SPAD = 4;   // smem row padding (floats); keeps float4 alignment

struct Opnd {
  const float* p;
  int ld;
  int kmajor;   // 1: elem(i,k) = p[i*ld + k]   0: elem(i,k) = p[k*ld + i]
};

// shared memory needed by gemm<TN> for TN <= 8
constexpr int kGemmSmemFloats = 2 * BK * (BM + SPAD) + 2 * BK * (16 * 8 + SPAD);

template <int IT>   // IT = tile extent along i (BM or BN)
struct TileLoader {
  static constexpr int NF = (IT * BK / 4 + kThreads - 1) / kThreads;   // float4 per thread
  float4 r[NF];

  __device__ __forceinline__ void load(const Opnd& o, int i0, int I, int k0, int K) {
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int f = threadIdx.x + j * kThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < IT * BK / 4) {
        if (o.kmajor) {
          const int i = i0 + (f >> 2), k = k0 + ((f & 3) << 2);
          if (i < I && k < K) {
            const float* src = o.p + (long long)i * o.ld + k;
            if (k + 3 < K) {
              v = *reinterpret_cast<const float4*>(src);
            } else {
              v.x = src[0];
              if (k + 1 < K) v.y = src[1];
              if (k + 2 < K) v.z = src[2];
            }
          }
        } else {
          const int k = k0 + f / (IT / 4), i = i0 + ((f % (IT / 4)) << 2);
          if (k < K && i < I) {
            const float* src = o.p + (long long)k * o.ld + i;
            if (i + 3 < I) {
              v = *reinterpret_cast<const float4*>(src);
            } else {
              v.x = src[0];
              if (i + 1 < I) v.y = src[1];
              if (i + 2 < I) v.z = src[2];
            }
          }
        }
      }
      r[j] = v;
    }
  }

  // smem tile layout: s[k][i], row stride IT + SPAD
  __device__ __forceinline__ void store(const Opnd& o, float* s) const {
    constexpr int LDS = IT + SPAD;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int f = threadIdx.x + j * kThreads;
      if (f < IT * BK / 4) {
        if (o.kmajor) {
          const int i = f >> 2, k = (f & 3) << 2;
          s[(k + 0) * LDS + i] = r[j].x;
          s[(k + 1) * LDS + i] = r[j].y;
          s[(k + 2) * LDS + i] = r[j].z;
          s[(k + 3) * LDS + i] = r[j].w;
        } else {
          const int k = f / (IT / 4), i = (f % (IT / 4)) << 2;
          *reinterpret_cast<float4*>(&s[k * LDS + i]) = r[j];
        }
      }
    }
  }
};

// All threads of the CTA must call with identical arguments.  `smem` >= kGemmSmemFloats floats.
// epi.row<TN>(m, nb, N, v) receives the TN values of row m at columns nb + 16*j (valid if < N).
template <int TN, class Epi>
__device__ __noinline__ void gemm(int M, int N, int K, Opnd A, Opnd B, Epi& epi, float* smem) {
  constexpr int BN = 16 * TN;
  constexpr int LDA = BM + SPAD, LDB = BN + SPAD;
  float* As = smem;
  float* Bs = smem + 2 * BK * LDA;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nk = (K + BK - 1) / BK;

  for (int m0 = 0; m0 < M; m0 += BM) {
    for (int n0 = 0; n0 < N; n0 += BN) {
      float acc[TM][TN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

      TileLoader<BM> la;
      TileLoader<BN> lb;
      la.load(A, m0, M, 0, K);
      lb.load(B, n0, N, 0, K);
      la.store(A, As);               // buffer 0 is free: the previous tile ended on a barrier
      lb.store(B, Bs);
      __syncthreads();

      for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
          la.load(A, m0, M, (kt + 1) * BK, K);
          lb.load(B, n0, N, (kt + 1) * BK, K);
        }
        const float* as = As + cur * BK * LDA + ty * TM;
        const float* bs = Bs + cur * BK * LDB + tx;
#pragma unroll 4
        for (int kk = 0; kk < BK; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(as + kk * LDA);
          const float4 a1 = *reinterpret_cast<const float4*>(as + kk * LDA + 4);
          const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float b[TN];
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j] = bs[kk * LDB + 16 * j];
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
          la.store(A, As + (cur ^ 1) * BK * LDA);
          lb.store(B, Bs + (cur ^ 1) * BK * LDB);
        }
        __syncthreads();
      }

      // Epilogue, one output row per call: the functor first issues ALL of its global loads for the
      // row (targets, activations, Adam state ...) and only then computes and stores, so a thread has
      // TN..3*TN independent loads in flight instead of one load -> use -> store chain per element.
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m < M) epi.template row<TN>(m, n0 + tx, N, acc[i]);
      }
    }
  }
  __syncthreads();   // epilogue stores visible to the CTA before the next stage reads them
}

// Pick the register-tile width whose padded N is smallest (ties -> wider tile).
template <class Epi>
__device__ __forceinline__ void gemm_auto(int M, int N, int K, Opnd A, Opnd B, Epi& epi, float* smem) {
  auto padded = [N](int bn) { return ((N + bn - 1) / bn) * bn; };
  const int p4 = padded(64), p5 = padded(80), p7 = padded(112), p8 = padded(128);
  int best = p8, tn = 8;
  if (p7 < best) { best = p7; tn = 7; }
  if (p5 < best) { best = p5; tn = 5; }
  if (p4 < best) { best = p4; tn = 4; }
  switch (tn) {
    case 4: gemm<4>(M, N, K, A, B, epi, smem); break;
    case 5: gemm<5>(M, N, K, A, B, epi, smem); break;
    case 7: gemm<7>(M, N, K, A, B, epi, smem); break;
    default: gemm<8>(M, N, K, A, B, epi, smem); break;
  }
}

}  // namespace nmb
