// Fused ensemble kernels: one CTA owns one ensemble member and runs, per minibatch step,
//   encoders -> latent fusion -> Philox reparameterisation -> decoders -> Gaussian-LL/-MSE + KL
//   -> backward -> Adam
// entirely inside a single launch for the whole ensemble (no host round trips between steps).
//
// Reference behaviour replaced (file:line in soz223/multi_modal_normative_modeling):
//   hot loop                     multimodal_kfold_train_cvae_supervised.py:177-199
//   Encoder/Decoder.forward      cVAE.py:161-172, 197-206
//   combine_latent + experts     cVAE.py:1144-1164, 986-1083
//   reparameterise               cVAE.py:1130-1133
//   calc_kl / compute_ll         cVAE.py:1138-1139, 14-15   (-MSE: ..._nmmlp.py:124-127)
//   loss_function_multimodal     cVAE.py:1187-1196
//   optimizer1 = Adam(enc+dec+alpha), torch defaults        cVAE.py:1111-1116
//   pred_recon (mean / sampled)  cVAE.py:549-555, 1198-1208
// The formulas are those of oracle/cvae_numpy.py.
#include "nmb_tc_gemm.cuh"
#include "nmb_internal.h"
#include "nmb_fusion.cuh"

namespace nmb {

struct StepCtx {
  const ArchDesc* a;
  MemberDev* mb;
  float* scratch;
  float* smem;      // FP32 engine: GEMM tiles
  tc::Ctx* tc;      // tensor-core engine context (null for the FP32 engine)
  float* red;       // reduction scratch (>= 16 floats)
  int rows, row0;
  const float* const* xin;   // per modality: base such that xin[m] + row0 * ldx is the minibatch (mb.xc, or staged rows)
  const int* yidx;           // head targets: y[yidx ? yidx[row0 + b] : row0 + b]
  long long step;   // global 0-based step index of this member
  unsigned flags;
  // Adam scalars of this step
  float step_size, bc2_sqrt, b1, b2, aeps;
};


// ---- epilogues ---------------------------------------------------------------------------
// row<TN>(m, nb, N, v): v[j] is C[m][nb + 16*j].  Loads first, then math + stores (see gemm()).
struct EpiHidden {     // h = leaky_relu(a) (bias already inside a via the ones column)
  float* dst; int ld; int nl;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    float* d = dst + (long long)m * ld + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (nb + 16 * j < N) d[16 * j] = (nl && v[j] <= 0.f) ? slope() * v[j] : v[j];
  }
  __device__ __forceinline__ float slope() const { return nl == 2 ? 0.f : kSlope; }   // nl: 1 = leaky_relu, 2 = ReLU (head)
  // tensor-core engine: v = C[m][n0 .. n0+16), first nvalid entries inside N
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float o[16];
    const float sl = slope();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (nl && v[j] <= 0.f) ? sl * v[j] : v[j];
    store16(dst + (long long)m * ld + n0, nvalid, o);
  }
};

struct EpiStore {
  float* dst; int ld;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    float* d = dst + (long long)m * ld + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (nb + 16 * j < N) d[16 * j] = v[j];
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float* d = dst + (long long)m * ld + n0;
    if ((ld & 3) == 0) { store16(d, nvalid, v); return; }
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < nvalid) d[j] = v[j];
  }
};

// x_recon epilogue: loss terms + d(total)/d(x_recon) in one pass.
struct EpiRecon {
  const float* x; int ldx;        // targets: packed rows (first D columns)
  const float* lam;               // logvar_out [D]
  float* dxh; int ld;             // out: gradient d(total)/d(x_recon)
  float* keep;                    // optional copy of x_recon (ld) for peek
  float inv_rows, inv_rows_d;     // 1/B, 1/(B*D)
  int gauss;
  float ll_acc;                   // per-thread partial of sum over elements
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    float xt[TN], l[TN];
    const float* xr = x + (long long)m * ldx + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const bool ok = nb + 16 * j < N;
      xt[j] = ok ? xr[16 * j] : 0.f;
      l[j] = (ok && gauss) ? lam[nb + 16 * j] : 0.f;
    }
    float* g = dxh + (long long)m * ld + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      if (nb + 16 * j < N) {
        const float r = xt[j] - v[j];
        if (gauss) {
          const float iv = __expf(-l[j]);            // 1 / sigma^2
          ll_acc += -0.5f * r * r * iv - 0.5f * l[j] - 0.5f * kLog2Pi;
          g[16 * j] = -r * iv * inv_rows;
        } else {
          ll_acc += -r * r;
          g[16 * j] = -2.f * r * inv_rows_d;
        }
        if (keep) keep[(long long)m * ld + nb + 16 * j] = v[j];
      }
    }
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float xt[16], l[16], g[16];
    load16(x + (long long)m * ldx + n0, nvalid, 0.f, xt);
    if (gauss) load16(lam + n0, nvalid, 0.f, l);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float r = xt[j] - v[j];
      float t, gg;
      if (gauss) {
        const float iv = __expf(-l[j]);
        t = -0.5f * r * r * iv - 0.5f * l[j] - 0.5f * kLog2Pi;
        gg = -r * iv * inv_rows;
      } else {
        t = -r * r;
        gg = -2.f * r * inv_rows_d;
      }
      if (j < nvalid) ll_acc += t;
      g[j] = gg;
    }
    store16(dxh + (long long)m * ld + n0, nvalid, g);
    if (keep) store16(keep + (long long)m * ld + n0, nvalid, v);
  }
};

struct EpiDgrad {      // d_pre = d_act * leaky_relu'(pre), sign recovered from the stored activation
  float* dst; int ld;
  const float* act; int ld_act; int nl;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    float h[TN];
    const float* a = act + (long long)m * ld_act + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j) h[j] = (nl && nb + 16 * j < N) ? a[16 * j] : 1.f;
    float* d = dst + (long long)m * ld + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (nb + 16 * j < N) d[16 * j] = h[j] <= 0.f ? (nl == 2 ? 0.f : kSlope) * v[j] : v[j];
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float h[16], o[16];
    if (nl) load16(act + (long long)m * ld_act + n0, nvalid, 1.f, h);
    const float sl = nl == 2 ? 0.f : kSlope;
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (nl && h[j] <= 0.f) ? sl * v[j] : v[j];
    store16(dst + (long long)m * ld + n0, nvalid, o);
  }
};

struct EpiDz {
  float* dz; int Z; int accumulate;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    float old[TN];
    float* p = dz + m * Z + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j) old[j] = (accumulate && nb + 16 * j < N) ? p[16 * j] : 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (nb + 16 * j < N) p[16 * j] = old[j] + v[j];
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float* p = dz + m * Z + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) p[j] = accumulate ? p[j] + v[j] : v[j];
  }
};

struct AdamCfg {
  float* p; float* m; float* v; float* g;
  float step_size, bc2_sqrt, b1, b2, eps;
  unsigned flags;
  __device__ __forceinline__ float update(float& m1, float& v1, float p0, float grad) const {
    m1 = b1 * m1 + (1.f - b1) * grad;
    v1 = b2 * v1 + (1.f - b2) * grad * grad;
    return p0 - step_size * (m1 / (sqrtf(v1) / bc2_sqrt + eps));
  }
  __device__ __forceinline__ void apply(long long idx, float grad) const {
    if (flags & NMB_TRAIN_WRITE_GRADS) g[idx] = grad;
    if (!(flags & NMB_TRAIN_NO_ADAM)) {
      float m1 = m[idx], v1 = v[idx];
      const float p1 = update(m1, v1, p[idx], grad);
      m[idx] = m1; v[idx] = v1; p[idx] = p1;
    }
  }
};

struct EpiWgradAdam {  // weight (+bias column) gradient fused with the optimiser update
  AdamCfg ad; long long off; int ld;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&gr)[TN]) {
    const long long base = off + (long long)m * ld + nb;
    if (ad.flags & NMB_TRAIN_WRITE_GRADS) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (nb + 16 * j < N) ad.g[base + 16 * j] = gr[j];
    }
    if (ad.flags & NMB_TRAIN_NO_ADAM) return;
    float p0[TN], m1[TN], v1[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const bool ok = nb + 16 * j < N;
      p0[j] = ok ? ad.p[base + 16 * j] : 0.f;
      m1[j] = ok ? ad.m[base + 16 * j] : 0.f;
      v1[j] = ok ? ad.v[base + 16 * j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      if (nb + 16 * j < N) {
        const float p1 = ad.update(m1[j], v1[j], p0[j], gr[j]);
        ad.m[base + 16 * j] = m1[j]; ad.v[base + 16 * j] = v1[j]; ad.p[base + 16 * j] = p1;
      }
    }
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&gr)[16]) {
    const long long base = off + (long long)m * ld + n0;
    if (ad.flags & NMB_TRAIN_WRITE_GRADS) store16(ad.g + base, nvalid, gr);
    if (ad.flags & NMB_TRAIN_NO_ADAM) return;
    float p0[16], m1[16], v1[16];
    load16(ad.p + base, nvalid, 0.f, p0);
    load16(ad.m + base, nvalid, 0.f, m1);
    load16(ad.v + base, nvalid, 0.f, v1);
#pragma unroll
    for (int j = 0; j < 16; ++j) p0[j] = ad.update(m1[j], v1[j], p0[j], gr[j]);
    store16(ad.m + base, nvalid, m1);
    store16(ad.v + base, nvalid, v1);
    store16(ad.p + base, nvalid, p0);
  }
};

__device__ __forceinline__ AdamCfg make_adam(const StepCtx& c) {
  AdamCfg ad;
  ad.p = c.mb->params; ad.m = c.mb->adam_m; ad.v = c.mb->adam_v; ad.g = c.mb->grads;
  ad.step_size = c.step_size; ad.bc2_sqrt = c.bc2_sqrt; ad.b1 = c.b1; ad.b2 = c.b2; ad.eps = c.aeps;
  ad.flags = c.flags;
  return ad;
}

// Engine dispatch: FP32 FFMA tiles (gemm_auto) or tcgen05 BF16x3 (tc::gemm).
template <bool TC, class Epi>
__device__ __forceinline__ void mm(const StepCtx& c, int M, int N, int K, Opnd A, Opnd B, Epi& epi) {
  if constexpr (TC) tc::gemm(M, N, K, A, B, epi, *c.tc);
  else gemm_auto(M, N, K, A, B, epi, c.smem);
}

// Constant-1 columns of the augmented activation buffers of this CTA's scratch slot.  Pad
// columns and stale rows are never read (every GEMM operand load is bounds-guarded).
__device__ void prepare_slot(const StepCtx& c) {
  const ArchDesc& a = *c.a;
  for (int m = 0; m < a.MD; ++m) {
    const ModDesc& q = a.mod[m];
    for (int b = threadIdx.x; b < kMaxBatch; b += kThreads) {
      for (int l = 0; l < a.L; ++l) {
        c.scratch[q.s_h[l] + (long long)b * q.ld_h[l] + a.hidden[l]] = 1.f;
        c.scratch[q.s_k[l] + (long long)b * q.ld_k[l] + a.hidden[a.L - 1 - l]] = 1.f;
      }
      c.scratch[q.s_g0 + (long long)b * q.ld_g0 + a.Z + a.C] = 1.f;
    }
  }
  if (a.head_kind) {
    for (int b = threadIdx.x; b < kMaxBatch; b += kThreads) {
      if (a.head_kind == NMB_HEAD_REGRESSION) c.scratch[a.s_R + (long long)b * a.ld_R + a.sumD] = 1.f;
      else c.scratch[a.s_zc + (long long)b * a.ld_zc + a.Z] = 1.f;
      for (int l = 0; l < a.HL; ++l) c.scratch[a.s_hh[l] + (long long)b * a.ld_hh[l] + a.head_w[l]] = 1.f;
    }
  }
  __syncthreads();
}

// ---- forward ---------------------------------------------------------------------------------
// Encoders of all modalities on rows [row0, row0+rows) of xc -> heads in scratch (s_mulv).
template <bool TC>
__device__ void encoders_forward(const StepCtx& c, const float* const* xc) {
  const ArchDesc& a = *c.a;
  const float* P = c.mb->params;
  for (int m = 0; m < a.M; ++m) {
    const ModDesc& q = a.mod[m];
    Opnd A{xc[m] + (long long)c.row0 * q.ldx, q.ldx, 1};
    for (int l = 0; l < a.L; ++l) {
      const LinDesc& w = q.enc[l];
      EpiHidden e{c.scratch + q.s_h[l], q.ld_h[l], a.non_linear};
      mm<TC>(c, c.rows, w.out, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
      A = Opnd{c.scratch + q.s_h[l], q.ld_h[l], 1};
    }
    EpiStore e{c.scratch + q.s_mulv, q.ld_mulv};
    mm<TC>(c, c.rows, q.head.out, q.head.in + 1, A, Opnd{P + q.head.off, q.head.ld, 1}, e);
  }
}

// Fusion + reparameterisation + KL.  Writes fused mu/logvar, eps, and the decoder inputs
// [z | c | 1] of every modality.  Returns sum over elements of the KL integrand (all threads).
//   eps_mode: 0 = Philox(stream_id), 1 = injected (eps_src [rows][Z]), 2 = zero (decode the mean),
//             3 = z itself is given in eps_src (decoders only; the heads are not read)
__device__ float latent_forward(const StepCtx& c, const float* const* xc, int eps_mode,
                                const float* eps_src, uint32_t stream_id, unsigned long long eps_step) {
  const ArchDesc& a = *c.a;
  const int Z = a.Z, M = a.M, n = c.rows * Z;
  float* S = c.scratch;
  float w[NMB_MAX_MOD];
  if (M > 1 && a.combine == NMB_COMBINE_GPOE) softmax_alpha(c.mb->params + a.alpha_off, M, w);
  float kl = 0.f;
  for (int g = threadIdx.x; g * 4 < n; g += kThreads) {
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (eps_mode == 0) philox_normal4(c.mb->seed, eps_step, stream_id, (uint32_t)g, nrm);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = g * 4 + j;
      if (e >= n) break;
      const int b = e / Z, z = e - b * Z;
      if (eps_mode == 3) {
        const float zz = eps_src[e];
        S[a.s_mub + e] = zz; S[a.s_lvb + e] = 0.f; S[a.s_eps + e] = 0.f;
        for (int m = 0; m < a.MD; ++m) S[a.mod[m].s_g0 + (long long)b * a.mod[m].ld_g0 + z] = zz;
        continue;
      }
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* h = S + a.mod[m].s_mulv + (long long)b * a.mod[m].ld_mulv;
        mu[m] = h[z]; lv[m] = h[Z + z];
      }
      bool clamped;
      const Fused f = a.family == NMB_FAMILY_MVTCAE ? fuse_forward_mvtcae(mu, lv, M, a.combine, w, &clamped)
                                                    : fuse_forward(mu, lv, M, a.combine, w);
      const float eps = eps_mode == 1 ? eps_src[e] : nrm[j];
      S[a.s_mub + e] = f.mu; S[a.s_lvb + e] = f.lv; S[a.s_eps + e] = eps;
      const float zz = f.mu + eps * expf(0.5f * f.lv);
      for (int m = 0; m < a.MD; ++m) S[a.mod[m].s_g0 + (long long)b * a.mod[m].ld_g0 + z] = zz;
      kl += -0.5f * (1.f + f.lv - f.mu * f.mu - expf(f.lv));
    }
  }
  // covariates into the decoder inputs (Decoder.forward: cat((z, c)), cVAE.py:199)
  for (int m = 0; m < a.MD; ++m) {
    const ModDesc& q = a.mod[m];
    const float* src = xc[m % M] + (long long)c.row0 * q.ldx + q.D;
    float* dst = S + q.s_g0 + Z;
    for (int e = threadIdx.x; e < c.rows * a.C; e += kThreads) {
      const int b = e / a.C, j = e - b * a.C;
      dst[(long long)b * q.ld_g0 + j] = src[(long long)b * q.ldx + j];
    }
  }
  __syncthreads();
  return kl;
}

// Decoder hidden layers of modality m; returns the operand feeding decoder_mean_layer.
template <bool TC>
__device__ Opnd decoder_hidden(const StepCtx& c, int m) {
  const ArchDesc& a = *c.a;
  const ModDesc& q = a.mod[m];
  const float* P = c.mb->params;
  Opnd A{c.scratch + q.s_g0, q.ld_g0, 1};
  for (int l = 0; l < a.L; ++l) {
    const LinDesc& w = q.dec[l];
    EpiHidden e{c.scratch + q.s_k[l], q.ld_k[l], a.non_linear};
    mm<TC>(c, c.rows, w.out, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
    A = Opnd{c.scratch + q.s_k[l], q.ld_k[l], 1};
  }
  return A;
}

// ---- supervised head (NMB_HEAD_REGRESSION, cVAE_multimodal_regression, cVAE.py:2211-2347) ---------------------
// Residual rows R = [x_0 - x_recon_0 | x_1 - x_recon_1 | ... | 1] from the targets and the kept reconstructions (:2321-2323).
__device__ void head_residuals(const StepCtx& c) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  for (int m = 0; m < a.M; ++m) {
    const ModDesc& q = a.mod[m];
    const float* x = c.xin[m] + (long long)c.row0 * q.ldx;
    const float* xr = S + q.s_xr;
    float* R = S + a.s_R + q.r_off;
    for (int e = threadIdx.x; e < c.rows * q.D; e += kThreads) {
      const int b = e / q.D, n = e - b * q.D;
      R[(long long)b * a.ld_R + n] = x[(long long)b * q.ldx + n] - xr[(long long)b * q.ld_xh + n];
    }
  }
  __syncthreads();
}

// regressor(R): Linear -> ReLU -> ... -> Linear(., 1) (:2248-2255); the prediction lands in s_pred[b * 4]
template <bool TC>
__device__ void head_forward(const StepCtx& c) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  const float* P = c.mb->params;
  Opnd A{S + a.s_R, a.ld_R, 1};
  for (int l = 0; l < a.HL; ++l) {
    const LinDesc& w = a.hd[l];
    EpiHidden e{S + a.s_hh[l], a.ld_hh[l], 2};
    mm<TC>(c, c.rows, w.out, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
    A = Opnd{S + a.s_hh[l], a.ld_hh[l], 1};
  }
  const LinDesc& w = a.hd[a.HL];
  EpiStore e{S + a.s_pred, 4};
  mm<TC>(c, c.rows, 1, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
}

// Forward + MSE (:2343) + backward + Adam of the head; leaves d(total)/dR in s_dR (head_fold).  Returns the
// regression loss (all threads).
template <bool TC>
__device__ float head_step(const StepCtx& c, const AdamCfg& ad) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  const float* P = c.mb->params;
  const int rows = c.rows;
  head_residuals(c);
  head_forward<TC>(c);
  float acc = 0.f;
  const float gsc = a.head_weight * 2.f / rows;
  for (int b = threadIdx.x; b < rows; b += kThreads) {
    const float yb = c.mb->y[c.yidx ? c.yidx[c.row0 + b] : c.row0 + b];
    const float r = S[a.s_pred + 4 * b] - yb;
    acc += r * r;
    S[a.s_dpred + 4 * b] = gsc * r;
  }
  const float loss = block_sum(acc, c.red) / rows;
  float* gbuf[2] = {S + a.s_ga, S + a.s_gb};
  const float* dy = S + a.s_dpred; int ld_dy = 4;
  int cur = 0;
  for (int l = a.HL; l >= 0; --l) {
    const LinDesc& w = a.hd[l];
    const float* in_act = l == 0 ? S + a.s_R : S + a.s_hh[l - 1];
    const int ld_in = l == 0 ? a.ld_R : a.ld_hh[l - 1];
    if (l > 0) {
      EpiDgrad eg{gbuf[cur], a.ld_g, in_act, ld_in, 2};
      mm<TC>(c, rows, w.in, w.out, Opnd{dy, ld_dy, 1}, Opnd{P + w.off, w.ld, 0}, eg);
    } else {
      EpiStore es{S + a.s_dR, a.ld_R};
      mm<TC>(c, rows, w.in, w.out, Opnd{dy, ld_dy, 1}, Opnd{P + w.off, w.ld, 0}, es);
    }
    EpiWgradAdam ew{ad, w.off, w.ld};
    mm<TC>(c, w.out, w.in + 1, rows, Opnd{dy, ld_dy, 0}, Opnd{in_act, ld_in, 0}, ew);
    if (l > 0) { dy = gbuf[cur]; ld_dy = a.ld_g; cur ^= 1; }
  }
  __syncthreads();
  return loss;
}

// d(total)/d(x_recon_m) += -d(total)/dR (R = x - x_recon).  After the logvar_out gradient, which recovers the residuals
// from the likelihood part of s_xh.
__device__ void head_fold(const StepCtx& c, int m) {
  const ArchDesc& a = *c.a;
  const ModDesc& q = a.mod[m];
  float* dxh = c.scratch + q.s_xh;
  const float* dR = c.scratch + a.s_dR + q.r_off;
  __syncthreads();
  for (int e = threadIdx.x; e < c.rows * q.D; e += kThreads) {
    const int b = e / q.D, n = e - b * q.D;
    dxh[(long long)b * q.ld_xh + n] -= dR[(long long)b * a.ld_R + n];
  }
  __syncthreads();
}

__device__ void e2e_fold(const StepCtx& c, int m);
__device__ void scale_dxh(const StepCtx& c, int m, float s);

// Backward of decoder set m (logvar_out, decoder_mean_layer, hidden layers down to dz) fused with Adam.
//   lam_scale: weight of the reconstruction term in the total (1; w_rec for the end-to-end model)
//   fold: fold another loss's d/d(x_recon) into s_xh after the logvar_out gradient (1 = regression head, 2 = contrastive)
template <bool TC>
__device__ void decoder_backward(const StepCtx& c, const AdamCfg& ad, int m, float lam_scale, bool dz_accumulate, int fold) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  float* P = c.mb->params;
  const int L = a.L, Z = a.Z, rows = c.rows;
  const float inv_rows = 1.f / rows;
  const int gauss = a.loss_kind == NMB_LOSS_GAUSS_LL;
  float* gbuf[2] = {S + a.s_ga, S + a.s_gb};
  {
    const ModDesc& q = a.mod[m];
    float* dxh = S + q.s_xh;
    // logvar_out: d/d lam_n = sum_b 0.5 (1 - r^2/sigma^2) / B, with r/sigma^2 = -B * dxh
    if (gauss) {
      for (int n = threadIdx.x; n < q.D; n += kThreads) {
        const float var = __expf(P[q.lam_off + n]);
        const float* col = dxh + n;
        float acc = 0.f;
        int b = 0;
        for (; b + 8 <= rows; b += 8) {
          float t[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = col[(long long)(b + u) * q.ld_xh];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const float w = t[u] * rows; acc += 0.5f * (1.f - w * w * var); }
        }
        for (; b < rows; ++b) { const float w = col[(long long)b * q.ld_xh] * rows; acc += 0.5f * (1.f - w * w * var); }
        ad.apply(q.lam_off + n, lam_scale * acc * inv_rows);
      }
    }
    if (fold == 1) head_fold(c, m);
    else if (fold == 2) e2e_fold(c, m);
    else if (fold == 3) scale_dxh(c, m, lam_scale);
    int cur = 0;
    // decoder_mean_layer
    {
      const LinDesc& w = q.outl;
      const float* act = S + q.s_k[L - 1]; const int ld_act = q.ld_k[L - 1];
      EpiDgrad eg{gbuf[cur], a.ld_g, act, ld_act, a.non_linear};
      mm<TC>(c, rows, w.in, w.out, Opnd{dxh, q.ld_xh, 1}, Opnd{P + w.off, w.ld, 0}, eg);
      EpiWgradAdam ew{ad, w.off, w.ld};
      mm<TC>(c, w.out, w.in + 1, rows, Opnd{dxh, q.ld_xh, 0}, Opnd{act, ld_act, 0}, ew);
    }
    for (int l = L - 1; l >= 0; --l) {
      const LinDesc& w = q.dec[l];
      const float* dy = gbuf[cur];
      const float* in_act = l == 0 ? S + q.s_g0 : S + q.s_k[l - 1];
      const int ld_in = l == 0 ? q.ld_g0 : q.ld_k[l - 1];
      if (l > 0) {
        EpiDgrad eg{gbuf[cur ^ 1], a.ld_g, in_act, ld_in, a.non_linear};
        mm<TC>(c, rows, w.in, w.out, Opnd{dy, a.ld_g, 1}, Opnd{P + w.off, w.ld, 0}, eg);
      } else {
        EpiDz ez{S + a.s_dz, Z, dz_accumulate};
        mm<TC>(c, rows, Z, w.out, Opnd{dy, a.ld_g, 1}, Opnd{P + w.off, w.ld, 0}, ez);
      }
      EpiWgradAdam ew{ad, w.off, w.ld};
      mm<TC>(c, w.out, w.in + 1, rows, Opnd{dy, a.ld_g, 0}, Opnd{in_act, ld_in, 0}, ew);
      cur ^= 1;
    }
  }
}

// klw: weight of the KL term in the total (M for sum_m (kl - ll_m); w_kl for the end-to-end model)
__device__ void latent_backward(const StepCtx& c, const AdamCfg& ad, float klw) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  float* P = c.mb->params;
  const int M = a.M, Z = a.Z, rows = c.rows;
  const float inv_rows = 1.f / rows;
  // latent backward: d mu_bar = dz + M mu/B ; d lv_bar = dz eps s/2 + M (e^lv - 1)/(2B)
  {
    float w[NMB_MAX_MOD], dw_acc[NMB_MAX_MOD];
    const bool gpoe = M > 1 && a.combine == NMB_COMBINE_GPOE;
    if (gpoe) softmax_alpha(P + a.alpha_off, M, w);
    for (int m = 0; m < M; ++m) dw_acc[m] = 0.f;
    for (int e = threadIdx.x; e < rows * Z; e += kThreads) {
      const int b = e / Z, z = e - b * Z;
      const float mub = S[a.s_mub + e], lvb = S[a.s_lvb + e], eps = S[a.s_eps + e], dz = S[a.s_dz + e];
      const float sd = expf(0.5f * lvb);
      const float dmu_bar = dz + klw * mub * inv_rows;
      const float dlv_bar = dz * eps * sd * 0.5f + klw * (expf(lvb) - 1.f) * 0.5f * inv_rows;
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD], dmu[NMB_MAX_MOD], dlv[NMB_MAX_MOD], dw[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* h = S + a.mod[m].s_mulv + (long long)b * a.mod[m].ld_mulv;
        mu[m] = h[z]; lv[m] = h[Z + z];
      }
      if (a.family == NMB_FAMILY_MVTCAE) {
        bool clamped;
        fuse_forward_mvtcae(mu, lv, M, a.combine, w, &clamped);
        fuse_backward_mvtcae(mu, lv, M, a.combine, w, clamped, dmu_bar, dlv_bar, dmu, dlv, gpoe ? dw : nullptr);
      } else {
        fuse_backward(mu, lv, M, a.combine, w, dmu_bar, dlv_bar, dmu, dlv, gpoe ? dw : nullptr);
      }
      for (int m = 0; m < M; ++m) {
        float* d = S + a.mod[m].s_dmulv + (long long)b * a.mod[m].ld_mulv;
        d[z] = dmu[m]; d[Z + z] = dlv[m];
        if (gpoe) dw_acc[m] += dw[m];
      }
    }
    if (gpoe) {   // softmax backward + Adam on alpha_m (cVAE.py:1154)
      float dw_tot[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) dw_tot[m] = block_sum(dw_acc[m], c.red);
      if (threadIdx.x == 0) {
        float dot = 0.f;
        for (int m = 0; m < M; ++m) dot += w[m] * dw_tot[m];
        for (int m = 0; m < M; ++m) ad.apply(a.alpha_off + m, w[m] * (dw_tot[m] - dot));
      }
    }
    __syncthreads();
  }
}

template <bool TC>
__device__ void encoder_backward(const StepCtx& c, const AdamCfg& ad, int m) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  float* P = c.mb->params;
  const int L = a.L, rows = c.rows;
  float* gbuf[2] = {S + a.s_ga, S + a.s_gb};
  {
    const ModDesc& q = a.mod[m];
    const float* dmulv = S + q.s_dmulv;
    int cur = 0;
    {
      const LinDesc& w = q.head;
      const float* act = S + q.s_h[L - 1]; const int ld_act = q.ld_h[L - 1];
      EpiDgrad eg{gbuf[cur], a.ld_g, act, ld_act, a.non_linear};
      mm<TC>(c, rows, w.in, w.out, Opnd{dmulv, q.ld_mulv, 1}, Opnd{P + w.off, w.ld, 0}, eg);
      EpiWgradAdam ew{ad, w.off, w.ld};
      mm<TC>(c, w.out, w.in + 1, rows, Opnd{dmulv, q.ld_mulv, 0}, Opnd{act, ld_act, 0}, ew);
    }
    for (int l = L - 1; l >= 0; --l) {
      const LinDesc& w = q.enc[l];
      const float* dy = gbuf[cur];
      const float* in_act = l == 0 ? c.xin[m] + (long long)c.row0 * q.ldx : S + q.s_h[l - 1];
      const int ld_in = l == 0 ? q.ldx : q.ld_h[l - 1];
      if (l > 0) {
        EpiDgrad eg{gbuf[cur ^ 1], a.ld_g, in_act, ld_in, a.non_linear};
        mm<TC>(c, rows, w.in, w.out, Opnd{dy, a.ld_g, 1}, Opnd{P + w.off, w.ld, 0}, eg);
      }
      EpiWgradAdam ew{ad, w.off, w.ld};
      mm<TC>(c, w.out, w.in + 1, rows, Opnd{dy, a.ld_g, 0}, Opnd{in_act, ld_in, 0}, ew);
      cur ^= 1;
    }
  }
}

// ---- end-to-end supervised model (NMB_HEAD_ENDTOEND, cVAE_multimodal_endtoend v2, cVAE.py:2004-2207) -------------
// Dropout keep flag of classifier unit (layer l, row b, unit j) at global step `step`: injected, or Philox stream 2.
__device__ __forceinline__ float e2e_keep(const StepCtx& c, int l, int b, int j, int col0, float p) {
  const MemberDev& mb = *c.mb;
  const ArchDesc& a = *c.a;
  if (p <= 0.f) return 1.f;
  if (mb.drop_keep)
    return mb.drop_keep[((c.step % mb.n_drop_steps) * mb.batch + b) * (long long)a.drop_w + col0 + j];
  const uint32_t e = (uint32_t)(b * a.drop_w + col0 + j);
  uint32_t w[4];
  philox4x32_10(e >> 2, (uint32_t)c.step, (uint32_t)((unsigned long long)c.step >> 32), 2u, (uint32_t)mb.seed,
                (uint32_t)(mb.seed >> 32), w);
  const float u = ((float)(w[e & 3] >> 9) + 0.5f) * 1.1920928955078125e-07f;
  return u >= p ? 1.f : 0.f;                       // bernoulli(1 - p)
}

// Classifier forward on [z | 1] (s_zc).  train: BatchNorm1d with batch statistics (+ running-statistics update,
// momentum 0.1, unbiased variance -- torch defaults), ReLU, Dropout(p); eval: running statistics, no dropout.
// Logits land in s_pred[b * 4 + {0, 1}].
template <bool TC>
__device__ void e2e_classifier_forward(const StepCtx& c, bool train) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  float* P = c.mb->params;
  const int rows = c.rows;
  const float p = a.hp[NMB_HP_DROPOUT];
  Opnd A{S + a.s_zc, a.ld_zc, 1};
  int col0 = 0;
  for (int l = 0; l < a.HL; ++l) {
    const LinDesc& w = a.hd[l];
    const int W = a.head_w[l], ld = a.ld_hh[l], bl = a.bn_ld[l];
    EpiStore e{S + a.s_xn[l], ld};
    mm<TC>(c, rows, w.out, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
    __syncthreads();
    float* bn = P + a.bn_off[l];                  // gamma | beta | running mean | running var | count
    float* xn = S + a.s_xn[l];
    float* istd = S + a.s_bn[l];
    for (int j = threadIdx.x; j < W; j += kThreads) {
      float mean, var;
      if (train) {
        float s1 = 0.f;
        for (int b = 0; b < rows; ++b) s1 += xn[(long long)b * ld + j];
        mean = s1 / rows;
        float s2 = 0.f;
        for (int b = 0; b < rows; ++b) { const float d = xn[(long long)b * ld + j] - mean; s2 += d * d; }
        var = s2 / rows;                           // biased: what normalises the batch
        if (!(c.flags & NMB_TRAIN_NO_STATS)) {
          bn[2 * bl + j] = 0.9f * bn[2 * bl + j] + 0.1f * mean;
          bn[3 * bl + j] = 0.9f * bn[3 * bl + j] + 0.1f * (rows > 1 ? s2 / (rows - 1) : var);
          if (j == 0) bn[4 * bl] += 1.f;
        }
      } else {
        mean = bn[2 * bl + j]; var = bn[3 * bl + j];
      }
      const float is = 1.f / sqrtf(var + 1e-5f);
      istd[j] = is;
      const float g = bn[j], be = bn[bl + j];
      for (int b = 0; b < rows; ++b) {
        const float v = (xn[(long long)b * ld + j] - mean) * is;
        const float y = g * v + be;
        float gate = y > 0.f ? 1.f : 0.f;
        if (train && p > 0.f) gate *= e2e_keep(c, l, b, j, col0, p) / (1.f - p);
        xn[(long long)b * ld + j] = v;
        S[a.s_gate[l] + (long long)b * ld + j] = gate;
        S[a.s_hh[l] + (long long)b * ld + j] = y * gate;
      }
    }
    __syncthreads();
    A = Opnd{S + a.s_hh[l], ld, 1};
    col0 += W;
  }
  const LinDesc& w = a.hd[a.HL];
  EpiStore e{S + a.s_pred, 4};
  mm<TC>(c, rows, 2, w.in + 1, A, Opnd{P + w.off, w.ld, 1}, e);
  __syncthreads();
}

__device__ void e2e_copy_z(const StepCtx& c) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  const ModDesc& q = a.mod[0];
  for (int e = threadIdx.x; e < c.rows * a.Z; e += kThreads) {
    const int b = e / a.Z, z = e - b * a.Z;
    S[a.s_zc + (long long)b * a.ld_zc + z] = S[q.s_g0 + (long long)b * q.ld_g0 + z];
  }
  __syncthreads();
}

// d(total)/d(x_recon) of decoder set m: w_rec x likelihood part (already in s_xh) + the contrastive hinge through the
// per-subject deviation mean_m mean_d (x - x_recon)^2 (cVAE.py:2150-2172); s_dev[b][2] holds w_con * sign_b / B.
__device__ void e2e_fold(const StepCtx& c, int m) {
  const ArchDesc& a = *c.a;
  const ModDesc& q = a.mod[m];
  float* S = c.scratch;
  float* dxh = S + q.s_xh;
  const float* xr = S + q.s_xr;
  const float* x = c.xin[m % a.M] + (long long)c.row0 * q.ldx;
  const float grp = m < a.M ? 1.f : -1.f;               // health decoders enter the hinge with +, disease with -
  const float k = 2.f / (q.D * a.M), w_rec = a.hp[NMB_HP_W_REC];
  __syncthreads();
  for (int e = threadIdx.x; e < c.rows * q.D; e += kThreads) {
    const int b = e / q.D, n = e - b * q.D;
    const float r = xr[(long long)b * q.ld_xh + n] - x[(long long)b * q.ldx + n];
    dxh[(long long)b * q.ld_xh + n] = w_rec * dxh[(long long)b * q.ld_xh + n] + grp * S[a.s_dev + 4 * b + 2] * k * r;
  }
  __syncthreads();
}

template <bool TC>
__device__ void train_step_e2e(StepCtx& c, const float* eps_src, float* loss_out) {
  const ArchDesc& a = *c.a;
  MemberDev& mb = *c.mb;
  float* S = c.scratch;
  float* P = mb.params;
  const int M = a.M, rows = c.rows;
  const float inv_rows = 1.f / rows;
  const float margin = a.hp[NMB_HP_MARGIN], w_con = a.hp[NMB_HP_W_CONTRASTIVE], w_kl = a.hp[NMB_HP_W_KL],
              w_rec = a.hp[NMB_HP_W_REC];

  // ---------------- forward (cVAE.py:2108-2125) ----------------
  encoders_forward<TC>(c, c.xin);
  float kl = latent_forward(c, c.xin, eps_src ? 1 : 0, eps_src, 0u, (unsigned long long)c.step);
  kl = block_sum(kl, c.red) * inv_rows;
  float rec[2] = {0.f, 0.f};
  for (int m = 0; m < a.MD; ++m) {
    const ModDesc& q = a.mod[m];
    Opnd A = decoder_hidden<TC>(c, m);
    EpiRecon e;
    e.x = c.xin[m % M] + (long long)c.row0 * q.ldx; e.ldx = q.ldx;
    e.lam = P + q.lam_off;
    e.dxh = S + q.s_xh; e.ld = q.ld_xh;
    e.keep = S + q.s_xr;
    e.inv_rows = inv_rows; e.inv_rows_d = inv_rows / q.D; e.gauss = 1; e.ll_acc = 0.f;
    mm<TC>(c, rows, q.D, q.outl.in + 1, A, Opnd{P + q.outl.off, q.outl.ld, 1}, e);
    rec[m < M ? 0 : 1] -= block_sum(e.ll_acc, c.red) * inv_rows;          // -log_prob(x).sum(1).mean() (:2131-2134)
  }
  __syncthreads();
  // per-subject deviations of the two decoder sets and the contrastive hinge (:2150-2172); one warp per row
  float con_acc = 0.f;
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < rows; b += kThreads / 32) {
      float dv[2] = {0.f, 0.f};
      for (int m = 0; m < a.MD; ++m) {
        const ModDesc& q = a.mod[m];
        const float* x = c.xin[m % M] + (long long)(c.row0 + b) * q.ldx;
        const float* xr = S + q.s_xr + (long long)b * q.ld_xh;
        float acc = 0.f;
        for (int n = lane; n < q.D; n += 32) { const float r = x[n] - xr[n]; acc += r * r; }
        dv[m < M ? 0 : 1] += warp_sum(acc) / q.D;
      }
      dv[0] /= M; dv[1] /= M;
      const float yb = mb.y[c.yidx ? c.yidx[c.row0 + b] : c.row0 + b];      // 0 = healthy, 1 = disease
      const float t = yb > 0.5f ? margin + dv[1] - dv[0] : margin + dv[0] - dv[1];
      float sgn = 0.f;
      if (t > 0.f) sgn = yb > 0.5f ? -1.f : 1.f;       // d hinge / d dev_health
      if (lane == 0) {
        S[a.s_dev + 4 * b + 0] = dv[0]; S[a.s_dev + 4 * b + 1] = dv[1];
        S[a.s_dev + 4 * b + 2] = w_con * sgn * inv_rows; S[a.s_dev + 4 * b + 3] = yb;
        con_acc += t > 0.f ? t : 0.f;
      }
    }
  }
  const float con = block_sum(con_acc, c.red) * inv_rows;
  // classifier on z, cross-entropy (:2118, :2178)
  e2e_copy_z(c);
  e2e_classifier_forward<TC>(c, true);
  float ce_acc = 0.f;
  for (int b = threadIdx.x; b < rows; b += kThreads) {
    const float l0 = S[a.s_pred + 4 * b], l1 = S[a.s_pred + 4 * b + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), den = e0 + e1;
    const int y = S[a.s_dev + 4 * b + 3] > 0.5f;
    ce_acc += -((y ? l1 : l0) - mx - logf(den));
    S[a.s_dpred + 4 * b] = (e0 / den - (y ? 0.f : 1.f)) * inv_rows;
    S[a.s_dpred + 4 * b + 1] = (e1 / den - (y ? 1.f : 0.f)) * inv_rows;
  }
  const float ce = block_sum(ce_acc, c.red) * inv_rows;
  if (loss_out && threadIdx.x == 0) {
    const float total = w_rec * (rec[0] + rec[1]) + w_kl * kl + ce + w_con * con;      // :2183
    loss_out[0] = total; loss_out[1] = kl; loss_out[2] = -(rec[0] + rec[1]);
    if (c.flags & (NMB_TRAIN_LOSS4 | NMB_TRAIN_LOSS8)) loss_out[3] = ce;
    if (c.flags & NMB_TRAIN_LOSS8) { loss_out[4] = rec[0]; loss_out[5] = rec[1]; loss_out[6] = con; loss_out[7] = 0.f; }
  }

  // ---------------- backward + Adam ----------------
  const AdamCfg ad = make_adam(c);
  {   // classifier
    float* gbuf[2] = {S + a.s_ga, S + a.s_gb};
    const float* dy = S + a.s_dpred; int ld_dy = 4;
    int cur = 0;
    for (int l = a.HL; l >= 0; --l) {
      const LinDesc& w = a.hd[l];
      const float* in_act = l == 0 ? S + a.s_zc : S + a.s_hh[l - 1];
      const int ld_in = l == 0 ? a.ld_zc : a.ld_hh[l - 1];
      if (l > 0) {
        EpiStore es{gbuf[cur], a.ld_g};
        mm<TC>(c, rows, w.in, w.out, Opnd{dy, ld_dy, 1}, Opnd{P + w.off, w.ld, 0}, es);
      } else {
        EpiDz ez{S + a.s_dz, a.Z, 0};
        mm<TC>(c, rows, a.Z, w.out, Opnd{dy, ld_dy, 1}, Opnd{P + w.off, w.ld, 0}, ez);
      }
      EpiWgradAdam ew{ad, w.off, w.ld};
      mm<TC>(c, w.out, w.in + 1, rows, Opnd{dy, ld_dy, 0}, Opnd{in_act, ld_in, 0}, ew);
      if (l > 0) {
        // through Dropout, ReLU and BatchNorm of hidden layer l - 1: gbuf[cur] = d h  ->  d(pre-BN activations), in place
        const int k = l - 1, W = a.head_w[k], ld = a.ld_hh[k], bl = a.bn_ld[k];
        float* g = gbuf[cur];
        const float* xn = S + a.s_xn[k];
        const float* gate = S + a.s_gate[k];
        const float* istd = S + a.s_bn[k];
        __syncthreads();
        for (int j = threadIdx.x; j < W; j += kThreads) {
          float sg = 0.f, sgx = 0.f;
          for (int b = 0; b < rows; ++b) {
            const float d = g[(long long)b * a.ld_g + j] * gate[(long long)b * ld + j];
            sg += d; sgx += d * xn[(long long)b * ld + j];
          }
          const float gam = P[a.bn_off[k] + j], is = istd[j];
          for (int b = 0; b < rows; ++b) {
            const float d = g[(long long)b * a.ld_g + j] * gate[(long long)b * ld + j];
            g[(long long)b * a.ld_g + j] = gam * is * (d - sg * inv_rows - xn[(long long)b * ld + j] * sgx * inv_rows);
          }
          ad.apply(a.bn_off[k] + j, sgx);               // d gamma
          ad.apply(a.bn_off[k] + bl + j, sg);           // d beta
        }
        __syncthreads();
        dy = g; ld_dy = a.ld_g; cur ^= 1;
      }
    }
    __syncthreads();
  }
  for (int m = 0; m < a.MD; ++m) decoder_backward<TC>(c, ad, m, w_rec, true, 2);     // dz accumulates onto the classifier's
  latent_backward(c, ad, w_kl);
  for (int m = 0; m < M; ++m) encoder_backward<TC>(c, ad, m);
}

// ---- DMVAE family (NMB_FAMILY_DMVAE: DMVAE / mmVAEPlus / WeightedDMVAE, cVAE.py:1454-1752, 1895-2002) ------------------
// x_recon = sigmoid(fc_out(.)) (VariationalDecoder, :1477-1480); ll_m = -0.5 sum_d (x - x_recon)^2 averaged over the batch
// (:1566); writes d(total)/d(pre-sigmoid) for total = ... - w_m ll_m.
struct EpiReconSig {
  const float* x; int ldx;
  float* dxh; int ld;
  float* keep;                    // x_recon [rows][ld] (always kept: peek / prediction)
  float gscale;                   // w_m / B
  float ll_acc;
  __device__ __forceinline__ void one(float xt, float a, float& g, float& xh) {
    xh = 1.f / (1.f + __expf(-a));
    const float r = xt - xh;
    ll_acc += -0.5f * r * r;
    g = -gscale * r * xh * (1.f - xh);
  }
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
    const float* xr = x + (long long)m * ldx + nb;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      if (nb + 16 * j < N) {
        float g, xh;
        one(xr[16 * j], v[j], g, xh);
        dxh[(long long)m * ld + nb + 16 * j] = g;
        keep[(long long)m * ld + nb + 16 * j] = xh;
      }
    }
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float xt[16], g[16], xh[16];
    load16(x + (long long)m * ldx + n0, nvalid, 0.f, xt);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float keep_acc = ll_acc;
      one(xt[j], v[j], g[j], xh[j]);
      if (j >= nvalid) ll_acc = keep_acc;
    }
    store16(dxh + (long long)m * ld + n0, nvalid, g);
    store16(keep + (long long)m * ld + n0, nvalid, xh);
  }
};

struct EpiStoreSig {   // prediction: x_recon = sigmoid(.)
  float* dst; int ld;
  template <int TN>
  __device__ __forceinline__ void row(int m, int nb, int N, const float (&v)[TN]) {
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (nb + 16 * j < N) dst[(long long)m * ld + nb + 16 * j] = 1.f / (1.f + __expf(-v[j]));
  }
  __device__ __forceinline__ void rowc(int m, int n0, int nvalid, const float (&v)[16]) {
    float* d = dst + (long long)m * ld + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j < nvalid) d[j] = 1.f / (1.f + __expf(-v[j]));
  }
};

// Heads -> decoder inputs [z_shared (Zc) | mu_private_m (S) | 1] (:1535-1554).  Shared part: ProductOfExperts2 over the
// modalities (:1482-1489), z = mu + eps * std; eps_mode as in latent_forward (eps_src: [rows][Z], the first Zc columns).  Returns this thread's
// part of sum over (row, shared dim) of the KL integrand.
__device__ float dmvae_latent_forward(const StepCtx& c, int eps_mode, const float* eps_src, uint32_t stream_id,
                                      unsigned long long eps_step) {
  const ArchDesc& a = *c.a;
  const int Z = a.Z, M = a.M, Sd = a.S, Zc = a.Zc, n = c.rows * Zc;
  float* S = c.scratch;
  float kl = 0.f;
  for (int g = threadIdx.x; g * 4 < n; g += kThreads) {
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (eps_mode == 0) philox_normal4(c.mb->seed, eps_step, stream_id, (uint32_t)g, nrm);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = g * 4 + j;
      if (e >= n) break;
      const int b = e / Zc, z = e - b * Zc;
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* h = S + a.mod[m].s_mulv + (long long)b * a.mod[m].ld_mulv;
        mu[m] = h[Sd + z]; lv[m] = h[Z + Sd + z];
      }
      const Fused f = fuse_forward(mu, lv, M, NMB_COMBINE_POE, nullptr);
      const float eps = eps_mode == 1 ? eps_src[b * Z + z] : (eps_mode == 2 ? 0.f : nrm[j]);   // injected: rows of `latent` entries
      S[a.s_mub + e] = f.mu; S[a.s_lvb + e] = f.lv; S[a.s_eps + e] = eps;
      const float zz = f.mu + eps * expf(0.5f * f.lv);
      for (int m = 0; m < M; ++m) S[a.mod[m].s_g0 + (long long)b * a.mod[m].ld_g0 + z] = zz;
      kl += -0.5f * (1.f + f.lv - f.mu * f.mu - expf(f.lv));
    }
  }
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    for (int e = threadIdx.x; e < c.rows * Sd; e += kThreads) {
      const int b = e / Sd, s = e - b * Sd;
      S[q.s_g0 + (long long)b * q.ld_g0 + Zc + s] = S[q.s_mulv + (long long)b * q.ld_mulv + s];
    }
  }
  __syncthreads();
  return kl;
}

template <bool TC>
__device__ void train_step_dmvae(StepCtx& c, const float* eps_src, float* loss_out) {
  const ArchDesc& a = *c.a;
  MemberDev& mb = *c.mb;
  float* S = c.scratch;
  float* P = mb.params;
  const int M = a.M, Z = a.Z, Sd = a.S, Zc = a.Zc, rows = c.rows;
  const float inv_rows = 1.f / rows;
  float w[NMB_MAX_MOD], ll[NMB_MAX_MOD], wsum = 0.f;
  for (int m = 0; m < M; ++m) { w[m] = a.weighted ? __ldcg(P + a.alpha_off + m) : 1.f; wsum += w[m]; }

  // ---------------- forward (:1535-1556) ----------------
  encoders_forward<TC>(c, c.xin);
  float kl = dmvae_latent_forward(c, eps_src ? 1 : 0, eps_src, 0u, (unsigned long long)c.step);
  kl = block_sum(kl, c.red) * inv_rows;
  float ll_tot = 0.f;
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    Opnd A = decoder_hidden<TC>(c, m);
    EpiReconSig e;
    e.x = c.xin[m] + (long long)c.row0 * q.ldx; e.ldx = q.ldx;
    e.dxh = S + q.s_xh; e.ld = q.ld_xh; e.keep = S + q.s_xr;
    e.gscale = w[m] * inv_rows; e.ll_acc = 0.f;
    mm<TC>(c, rows, q.D, q.outl.in + 1, A, Opnd{P + q.outl.off, q.outl.ld, 1}, e);
    ll[m] = block_sum(e.ll_acc, c.red) * inv_rows;
    ll_tot += w[m] * ll[m];
  }
  if (loss_out && threadIdx.x == 0) {      // (:1562-1572; WeightedDMVAE :1693-1707 has beta = 1)
    loss_out[0] = a.beta * wsum * kl - ll_tot; loss_out[1] = wsum * kl; loss_out[2] = ll_tot;
  }

  // ---------------- backward + Adam ----------------
  const AdamCfg ad = make_adam(c);
  for (int e = threadIdx.x; e < rows * Z; e += kThreads) S[a.s_dzs + e] = 0.f;
  __syncthreads();
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    decoder_backward<TC>(c, ad, m, 1.f, false, 0);          // d(total)/d[z_shared | mu_private_m] -> s_dz
    __syncthreads();
    for (int e = threadIdx.x; e < rows * Z; e += kThreads) {
      const int b = e / Z, j = e - b * Z;
      const float g = S[a.s_dz + e];
      if (j < Zc) S[a.s_dzs + e] += g;
      else {                                               // the private mean feeds only its own decoder; its logvar feeds nothing
        float* d = S + q.s_dmulv + (long long)b * q.ld_mulv;
        d[j - Zc] = g; d[Z + j - Zc] = 0.f;
      }
    }
    if (a.weighted && threadIdx.x == 0) ad.apply(a.alpha_off + m, kl - ll[m]);      // d total / d weights[m]
    __syncthreads();
  }
  {   // shared latent: reparameterisation, KL (weight beta * sum_m w_m) and the product of experts
    const float klw = a.beta * wsum;
    for (int e = threadIdx.x; e < rows * Zc; e += kThreads) {
      const int b = e / Zc, z = e - b * Zc;
      const float mub = S[a.s_mub + e], lvb = S[a.s_lvb + e], eps = S[a.s_eps + e], dz = S[a.s_dzs + b * Z + z];
      const float sd = expf(0.5f * lvb);
      const float dmu_bar = dz + klw * mub * inv_rows;
      const float dlv_bar = dz * eps * sd * 0.5f + klw * (expf(lvb) - 1.f) * 0.5f * inv_rows;
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD], dmu[NMB_MAX_MOD], dlv[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* h = S + a.mod[m].s_mulv + (long long)b * a.mod[m].ld_mulv;
        mu[m] = h[Sd + z]; lv[m] = h[Z + Sd + z];
      }
      fuse_backward(mu, lv, M, NMB_COMBINE_POE, nullptr, dmu_bar, dlv_bar, dmu, dlv, nullptr);
      for (int m = 0; m < M; ++m) {
        float* d = S + a.mod[m].s_dmulv + (long long)b * a.mod[m].ld_mulv;
        d[Sd + z] = dmu[m]; d[Z + Sd + z] = dlv[m];
      }
    }
    __syncthreads();
  }
  for (int m = 0; m < M; ++m) encoder_backward<TC>(c, ad, m);
}

// ---- mvtCAE (NMB_FAMILY_MVTCAE, cVAE.py:1754-1893) -----------------------------------------------------------------
// total_correlation as written (:1846-1853): log_qz_xi = lse - mean(lse) of a scalar = 0, so
//   tc = - sum_i mean_m logsumexp_b mu_m[b, i]   (mu_m = the per-modality encoder means, qz_xs).
// grad_coef == 0: returns tc (all threads).  grad_coef != 0: adds grad_coef * d tc / d mu_m[b, i] = -grad_coef / M *
// softmax_b(mu_m[:, i])[b] to the head gradients (s_dmulv).
__device__ float mvtcae_tc(const StepCtx& c, float grad_coef) {
  const ArchDesc& a = *c.a;
  float* S = c.scratch;
  const int M = a.M, Z = a.Z, rows = c.rows;
  float acc = 0.f;
  for (int p = threadIdx.x; p < M * Z; p += kThreads) {
    const int m = p / Z, i = p - m * Z;
    const float* mu = S + a.mod[m].s_mulv + i;
    const int ld = a.mod[m].ld_mulv;
    float mx = mu[0];
    for (int b = 1; b < rows; ++b) mx = fmaxf(mx, mu[(long long)b * ld]);
    float se = 0.f;
    for (int b = 0; b < rows; ++b) se += expf(mu[(long long)b * ld] - mx);
    acc -= (mx + logf(se)) / M;
    if (grad_coef != 0.f) {
      float* d = S + a.mod[m].s_dmulv + i;
      for (int b = 0; b < rows; ++b) d[(long long)b * ld] += -grad_coef / M * expf(mu[(long long)b * ld] - mx) / se;
    }
  }
  if (grad_coef != 0.f) { __syncthreads(); return 0.f; }
  return block_sum(acc, c.red);
}

// d(total)/d(x_recon) of mvtCAE: the log-likelihood enters the total with weight +1e-5 instead of -1 (:1862)
__device__ void scale_dxh(const StepCtx& c, int m, float s) {
  const ModDesc& q = c.a->mod[m];
  float* dxh = c.scratch + q.s_xh;
  __syncthreads();
  for (int e = threadIdx.x; e < c.rows * q.D; e += kThreads) {
    const int b = e / q.D, n = e - b * q.D;
    dxh[(long long)b * q.ld_xh + n] *= s;
  }
  __syncthreads();
}

// ---- one training step ---------------------------------------------------------------------
template <bool TC>
__device__ void train_step(StepCtx& c, const float* eps_src, float* loss_out) {
  const ArchDesc& a = *c.a;
  MemberDev& mb = *c.mb;
  float* S = c.scratch;
  float* P = mb.params;
  const int M = a.M, L = a.L, Z = a.Z, rows = c.rows;
  const float inv_rows = 1.f / rows;
  const int gauss = a.loss_kind == NMB_LOSS_GAUSS_LL;

  // ---------------- forward ----------------
  encoders_forward<TC>(c, c.xin);
  float kl = latent_forward(c, c.xin, eps_src ? 1 : 0, eps_src, 0u, (unsigned long long)c.step);
  kl = block_sum(kl, c.red) * inv_rows;
  float ll_sum = 0.f;
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    Opnd A = decoder_hidden<TC>(c, m);
    EpiRecon e;
    e.x = c.xin[m] + (long long)c.row0 * q.ldx; e.ldx = q.ldx;
    e.lam = P + q.lam_off;
    e.dxh = S + q.s_xh; e.ld = q.ld_xh;
    e.keep = ((c.flags & NMB_TRAIN_KEEP_ACTS) || a.head_kind) ? S + q.s_xr : nullptr;
    e.inv_rows = inv_rows; e.inv_rows_d = inv_rows / q.D; e.gauss = gauss; e.ll_acc = 0.f;
    mm<TC>(c, rows, q.D, q.outl.in + 1, A, Opnd{P + q.outl.off, q.outl.ld, 1}, e);
    const float s = block_sum(e.ll_acc, c.red);
    ll_sum += gauss ? s * inv_rows : s * inv_rows / q.D;
  }
  // ---------------- backward + Adam ----------------
  const AdamCfg ad = make_adam(c);
  float head_loss = 0.f;
  if (a.head_kind) head_loss = head_step<TC>(c, ad);
  const bool mvt = a.family == NMB_FAMILY_MVTCAE;
  const float tc = mvt ? mvtcae_tc(c, 0.f) : 0.f;
  if (loss_out && threadIdx.x == 0) {
    loss_out[0] = M * kl - ll_sum + a.head_weight * head_loss; loss_out[1] = M * kl; loss_out[2] = ll_sum;   // cVAE.py:1187-1196
    if (mvt) loss_out[0] = M * kl + 1e-5f * ll_sum + M * a.beta * tc;                                        // cVAE.py:1858-1868
    if (c.flags & NMB_TRAIN_LOSS4) loss_out[3] = mvt ? M * tc : head_loss;                                   // cVAE.py:2343-2345 / losses['tc']
  }
  for (int m = 0; m < M; ++m) decoder_backward<TC>(c, ad, m, mvt ? -1e-5f : 1.f, m > 0, mvt ? 3 : (a.head_kind ? 1 : 0));
  latent_backward(c, ad, (float)M);
  if (mvt) mvtcae_tc(c, M * a.beta);
  for (int m = 0; m < M; ++m) encoder_backward<TC>(c, ad, m);
}

// Members trained through shuffling loaders: copy the minibatch rows each modality's permutation selects into the
// staging buffers and point xin at them (xin[m] + row0 * ldx = staged row 0).
__device__ void stage_rows(StepCtx& c, long long epoch, const float** xin) {
  const ArchDesc& a = *c.a;
  const MemberDev& mb = *c.mb;
  __syncthreads();
  for (int m = 0; m < a.M; ++m) {
    const ModDesc& q = a.mod[m];
    const int* ord = mb.row_order + (epoch * a.M + m) * (long long)mb.n_rows + c.row0;
    float* dst = c.scratch + q.s_in;
    const int w4 = q.ldx / 4;
    for (int e = threadIdx.x; e < c.rows * w4; e += kThreads) {
      const int b = e / w4, j = e - b * w4;
      reinterpret_cast<float4*>(dst + (long long)b * q.ldx)[j] =
          reinterpret_cast<const float4*>(mb.xc[m] + (long long)ord[b] * q.ldx)[j];
    }
    if (threadIdx.x == 0) xin[m] = dst - (long long)c.row0 * q.ldx;
  }
  c.yidx = mb.row_order + epoch * a.M * (long long)mb.n_rows;       // the target follows modality 0's loader
  __syncthreads();
}

// ---- kernels -------------------------------------------------------------------------------
// One CTA = one ensemble member at a time; TC selects the GEMM engine of every dense stage.
template <bool TC>
__device__ __forceinline__ void train_body(const TrainLaunch& t, float* smem_f, tc::Ctx* tcx, float* red, int* s_member) {
  for (;;) {
    int mi;
    if ((int)gridDim.x >= t.n_members) {
      mi = blockIdx.x;
    } else {
      if (threadIdx.x == 0) *s_member = atomicAdd(t.work_counter, 1);
      __syncthreads();
      mi = *s_member;
      __syncthreads();
      if (mi >= t.n_members) return;
      mi = t.order[mi];
    }
    if (mi >= t.n_members) return;
    MemberDev& mb = t.members[mi];
    const ArchDesc& a = t.archs[mb.arch_idx];
    StepCtx c;
    c.a = &a; c.mb = &mb; c.smem = smem_f; c.tc = tcx; c.red = red; c.flags = t.flags;
    c.scratch = t.scratch + (long long)blockIdx.x * t.slot_floats;
    c.b1 = mb.beta1; c.b2 = mb.beta2; c.aeps = mb.adam_eps;
    prepare_slot(c);
    __shared__ const float* s_xin[NMB_MAX_MOD];
    if (threadIdx.x < NMB_MAX_MOD) s_xin[threadIdx.x] = mb.xc[threadIdx.x];
    c.xin = s_xin; c.yidx = nullptr;
    __syncthreads();
    const int spe = (mb.n_rows + mb.batch - 1) / mb.batch;   // steps per epoch
    const long long s0 = mb.steps_done;
    int rows = 0;
    const long long ns = member_steps(t, mb);
    for (long long i = 0; i < ns; ++i) {
      const long long s = s0 + i;
      const int pos = (int)(s % spe);
      c.step = s;
      c.row0 = pos * mb.batch;
      c.rows = rows = min(mb.batch, mb.n_rows - c.row0);
      if (mb.row_order) stage_rows(c, s / spe, s_xin);      // this epoch's per-modality permutations
      const double tt = (double)(s + 1);
      const float lr = mb.lr_steps ? mb.lr_steps[s] : mb.lr;
      c.step_size = (float)((double)lr / (1.0 - pow((double)mb.beta1, tt)));
      c.bc2_sqrt = (float)sqrt(1.0 - pow((double)mb.beta2, tt));
      const float* eps = t.eps_override
          ? t.eps_override + ((long long)mi * t.stride_steps + i) * mb.batch * a.Z : nullptr;
      float* lo = t.loss_out ? t.loss_out + ((long long)mi * t.stride_steps + i) * ((t.flags & NMB_TRAIN_LOSS8) ? 8 : (t.flags & NMB_TRAIN_LOSS4) ? 4 : 3) : nullptr;
      if (a.family == NMB_FAMILY_DMVAE) train_step_dmvae<TC>(c, eps, lo);
      else if (a.head_kind == NMB_HEAD_ENDTOEND) train_step_e2e<TC>(c, eps, lo);
      else train_step<TC>(c, eps, lo);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mb.steps_done = s0 + ns;
      mb.last_rows = rows;
      mb.last_slot = blockIdx.x;
    }
    if ((int)gridDim.x >= t.n_members) return;
  }
}

// FP32 FFMA engine: bit-stable trajectories (NMB_TRAIN_FP32).
__global__ void __launch_bounds__(kThreads, 2) train_kernel(TrainLaunch t) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[16];
  __shared__ int s_member;
  train_body<false>(t, smem, nullptr, red, &s_member);
}

// tcgen05 engine (default): BF16x3 split products, FP32 accumulation in TMEM.
__global__ void __launch_bounds__(kThreads, 2) train_tc_kernel(TrainLaunch t) {
  extern __shared__ __align__(128) unsigned char smem_tc[];
  __shared__ float red[16];
  __shared__ int s_member;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_mbar;
  if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, tc::kTmemCols);
  if (threadIdx.x == 0) tc::mbar_init(&s_mbar, 1);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  tc::Ctx tcx;
  tcx.smem = smem_tc; tcx.mbar = &s_mbar; tcx.tmem_base = s_tmem; tcx.phase = 0;
  train_body<true>(t, nullptr, &tcx, red, &s_member);
  tc::fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_free(tcx.tmem_base, tc::kTmemCols);
}

// Test hook: one CTA computes C[m][n] = sum_k A(m,k) B(n,k) with the tensor-core engine.
__global__ void __launch_bounds__(kThreads, 2) debug_tc_gemm_kernel(Opnd A, Opnd B, float* C, int ldc, int M, int N, int K) {
  extern __shared__ __align__(128) unsigned char smem_tc[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_mbar;
  if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, tc::kTmemCols);
  if (threadIdx.x == 0) tc::mbar_init(&s_mbar, 1);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  tc::Ctx tcx;
  tcx.smem = smem_tc; tcx.mbar = &s_mbar; tcx.tmem_base = s_tmem; tcx.phase = 0;
  EpiStore e{C, ldc};
  tc::gemm(M, N, K, A, B, e, tcx);
  tc::fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_free(tcx.tmem_base, tc::kTmemCols);
}

// Test-time reconstruction.  Work item = (member, tile of up to 256 rows).
template <bool TC>
__device__ __forceinline__ void recon_body(const ReconLaunch& t, float* smem_f, tc::Ctx* tcx, float* red) {
  for (int it = blockIdx.x; it < t.n_items; it += gridDim.x) {
    const ReconItem item = t.items[it];
    MemberDev& mb = t.members[item.member];
    const ArchDesc& a = t.archs[mb.arch_idx];
    StepCtx c;
    c.a = &a; c.mb = &mb; c.smem = smem_f; c.tc = tcx; c.red = red; c.flags = 0;
    c.scratch = t.scratch + (long long)blockIdx.x * t.slot_floats;
    c.rows = item.rows; c.row0 = item.row0; c.step = 0;
    const float* const* xc = t.xc + (long long)item.member * NMB_MAX_MOD;
    c.xin = xc; c.yidx = nullptr;
    float* head_out = (a.head_kind && t.head_out) ? t.head_out[item.member] : nullptr;
    prepare_slot(c);
    if (t.mode != NMB_RECON_GIVEN_Z) encoders_forward<TC>(c, xc);
    const float* eps = (t.mode != NMB_RECON_MEAN && t.eps && t.eps[item.member])
        ? t.eps[item.member] + (long long)item.row0 * a.Z : nullptr;
    const int eps_mode = t.mode == NMB_RECON_GIVEN_Z ? 3 : (t.mode == NMB_RECON_MEAN ? 2 : (eps ? 1 : 0));
    // Philox test stream: counter "step" = row tile index so that tiles draw disjoint numbers
    if (a.family == NMB_FAMILY_DMVAE) {      // pred_recon of the DMVAE family (cVAE.py:1574-1596): shared z sampled, private means
      dmvae_latent_forward(c, eps_mode == 3 ? 2 : eps_mode, eps, 1u, (unsigned long long)(item.row0 / kMaxBatch));
      for (int m = 0; m < a.M; ++m) {
        const ModDesc& q = a.mod[m];
        float* out = t.xhat[(long long)item.member * NMB_MAX_MOD + m];
        if (!out) continue;
        Opnd A = decoder_hidden<TC>(c, m);
        EpiStoreSig e{out + (long long)item.row0 * q.D, q.D};
        mm<TC>(c, item.rows, q.D, q.outl.in + 1, A, Opnd{mb.params + q.outl.off, q.outl.ld, 1}, e);
      }
      __syncthreads();
      continue;
    }
    latent_forward(c, xc, eps_mode, eps, 1u, (unsigned long long)(item.row0 / kMaxBatch));
    if (head_out && a.head_kind == NMB_HEAD_ENDTOEND) {
      // predict(): logits = classifier(z) in eval mode (running statistics, no dropout); the reference calls it on the
      // fused mean (mode MEAN, cVAE.py:2198-2203).  Decoders run only if reconstructions were asked for.
      e2e_copy_z(c);
      e2e_classifier_forward<TC>(c, false);
      for (int e2 = threadIdx.x; e2 < item.rows * 2; e2 += kThreads)
        head_out[(long long)(item.row0 + (e2 >> 1)) * 2 + (e2 & 1)] = c.scratch[a.s_pred + 4 * (e2 >> 1) + (e2 & 1)];
      head_out = nullptr;
    }
    if (t.mode != NMB_RECON_GIVEN_Z && t.mu && t.mu[item.member]) {
      float* mu = t.mu[item.member] + (long long)item.row0 * a.Z;
      for (int e = threadIdx.x; e < item.rows * a.Z; e += kThreads) mu[e] = c.scratch[a.s_mub + e];
    }
    if (t.mode != NMB_RECON_GIVEN_Z && t.logvar && t.logvar[item.member]) {
      float* lv = t.logvar[item.member] + (long long)item.row0 * a.Z;
      for (int e = threadIdx.x; e < item.rows * a.Z; e += kThreads) lv[e] = c.scratch[a.s_lvb + e];
    }
    for (int m = 0; m < a.MD; ++m) {
      const ModDesc& q = a.mod[m];
      float* out = t.xhat[(long long)item.member * NMB_MAX_MOD + m];
      if (!out && !head_out) continue;
      Opnd A = decoder_hidden<TC>(c, m);
      if (head_out) {      // the head reads the reconstructions from scratch
        EpiStore e{c.scratch + q.s_xr, q.ld_xh};
        mm<TC>(c, item.rows, q.D, q.outl.in + 1, A, Opnd{mb.params + q.outl.off, q.outl.ld, 1}, e);
        __syncthreads();
        if (out)
          for (int e2 = threadIdx.x; e2 < item.rows * q.D; e2 += kThreads) {
            const int b = e2 / q.D, n = e2 - b * q.D;
            out[(long long)(item.row0 + b) * q.D + n] = c.scratch[q.s_xr + (long long)b * q.ld_xh + n];
          }
        continue;
      }
      if (!out) continue;
      EpiStore e{out + (long long)item.row0 * q.D, q.D};
      mm<TC>(c, item.rows, q.D, q.outl.in + 1, A, Opnd{mb.params + q.outl.off, q.outl.ld, 1}, e);
    }
    if (head_out) {        // fi_pred = regressor(cat(x_m - x_recon_m.loc)) (cVAE.py:2321-2325)
      __syncthreads();
      head_residuals(c);
      head_forward<TC>(c);
      __syncthreads();
      for (int b = threadIdx.x; b < item.rows; b += kThreads) head_out[item.row0 + b] = c.scratch[a.s_pred + 4 * b];
    }
    __syncthreads();
  }
}

// FP32 engine (NMB_RECON_FP32 bit of `mode`)
__global__ void __launch_bounds__(kThreads, 2) recon_kernel(ReconLaunch t) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[16];
  recon_body<false>(t, smem, nullptr, red);
}

// tcgen05 engine (default): BF16x3 split products, FP32 accumulation in TMEM
__global__ void __launch_bounds__(kThreads, 2) recon_tc_kernel(ReconLaunch t) {
  extern __shared__ __align__(128) unsigned char smem_tc[];
  __shared__ float red[16];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_mbar;
  if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, tc::kTmemCols);
  if (threadIdx.x == 0) tc::mbar_init(&s_mbar, 1);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  tc::Ctx tcx;
  tcx.smem = smem_tc; tcx.mbar = &s_mbar; tcx.tmem_base = s_tmem; tcx.phase = 0;
  recon_body<true>(t, nullptr, &tcx, red);
  tc::fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_free(tcx.tmem_base, tc::kTmemCols);
}

constexpr size_t kGemmSmemBytes = kGemmSmemFloats * sizeof(float);

cudaError_t configure_kernels() {
  cudaError_t e = cudaFuncSetAttribute(train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(train_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(debug_tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(recon_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(recon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes);
}

cudaError_t launch_train(const TrainLaunch& t, cudaStream_t st) {
  const int grid = t.n_members < t.n_slots ? t.n_members : t.n_slots;
  if (grid <= 0 || t.n_steps <= 0) return cudaSuccess;
  if (grid < t.n_members) {
    cudaError_t e = cudaMemsetAsync(t.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  if (t.flags & NMB_TRAIN_FP32) train_kernel<<<grid, kThreads, kGemmSmemBytes, st>>>(t);
  else train_tc_kernel<<<grid, kThreads, tc::kSmemBytes, st>>>(t);
  return cudaGetLastError();
}

cudaError_t launch_debug_tc_gemm(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor,
                                 float* C, int ldc, int M, int N, int K, cudaStream_t st) {
  debug_tc_gemm_kernel<<<1, kThreads, tc::kSmemBytes, st>>>(Opnd{A, lda, a_kmajor}, Opnd{B, ldb, b_kmajor}, C, ldc, M, N, K);
  return cudaGetLastError();
}

cudaError_t launch_recon(const ReconLaunch& t, cudaStream_t st) {
  const int grid = t.n_items < t.n_slots ? t.n_items : t.n_slots;
  if (grid <= 0) return cudaSuccess;
  if (t.fp32) recon_kernel<<<grid, kThreads, kGemmSmemBytes, st>>>(t);
  else recon_tc_kernel<<<grid, kThreads, tc::kSmemBytes, st>>>(t);
  return cudaGetLastError();
}

}  // namespace nmb
