// Shared host/device definitions for libnmb: packed-parameter layout, scratch layout,
// Philox eps stream, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nmb.h"

namespace nmb {

constexpr int kThreads = 256;      // threads per CTA of every GEMM-bearing kernel
constexpr int kMaxBatch = 256;     // minibatch rows a CTA processes per step / row tile
constexpr float kSlope = 0.01f;    // F.leaky_relu default (cVAE.py:167)
constexpr float kLog2Pi = 1.8378770664093453f;

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

// One linear layer stored augmented: W_aug[out][ld], ld = round4(in + 1), column `in` = bias.
struct LinDesc {
  int in, out, ld;
  long long off;   // float offset in the packed parameter buffer
};

struct ModDesc {
  int D, ldx;                            // ROI count; stride of packed dataset rows [x | c | 1]
  LinDesc enc[NMB_MAX_HIDDEN], head, dec[NMB_MAX_HIDDEN], outl;
  long long lam_off;                     // logvar_out [D]
  // per-CTA scratch (float offsets).  Activations are stored post-nonlinearity with a
  // constant-1 column appended (index = width) so that biases ride through the GEMMs.
  long long s_h[NMB_MAX_HIDDEN];  int ld_h[NMB_MAX_HIDDEN];   // encoder hidden
  long long s_k[NMB_MAX_HIDDEN];  int ld_k[NMB_MAX_HIDDEN];   // decoder hidden
  long long s_mulv, s_dmulv;      int ld_mulv;                // [mu | logvar] heads and their grads
  long long s_g0;                 int ld_g0;                  // decoder input [z | c | 1]
  long long s_xh;                 int ld_xh;                  // d(total)/d(x_recon)
  long long s_xr;                                             // x_recon kept for nmb_ensemble_peek
  long long s_in;                                             // staged minibatch rows (members with row_order)
  int r_off;                                                  // first column of this modality in the head's residual row
};

struct ArchDesc {
  int M, L, Z, C, combine, loss_kind, non_linear;
  int MD;                                // decoder sets: M, or 2 M with NMB_HEAD_ENDTOEND (mod[M + m] = disease decoder of m)
  int hidden[NMB_MAX_HIDDEN];
  ModDesc mod[NMB_MAX_MOD];
  long long alpha_off;
  long long n_params;
  long long s_mub, s_lvb, s_eps, s_dz;   // fused mu, fused logvar, eps, d(total)/dz   [B][Z]
  long long s_ga, s_gb; int ld_g;        // gradient ping-pong buffers [B][ld_g]
  // supervised head (NMB_HEAD_REGRESSION): linears hd[0..HL], input = residual row [x_m - x_recon_m ... | 1]
  int head_kind, HL, sumD;
  float head_weight;
  int head_w[NMB_MAX_HEAD];
  LinDesc hd[NMB_MAX_HEAD + 1];
  long long s_R, s_dR; int ld_R;                              // residual rows and their gradient
  long long s_hh[NMB_MAX_HEAD]; int ld_hh[NMB_MAX_HEAD];      // head hidden activations (post-ReLU | 1)
  long long s_pred, s_dpred;                                  // [B][4]: prediction / logits, d(total)/d(.)
  // NMB_HEAD_ENDTOEND: classifier on z with BatchNorm + dropout after every hidden layer
  float hp[6];                                                // NMB_HP_*
  long long bn_off[NMB_MAX_HEAD]; int bn_ld[NMB_MAX_HEAD];    // 5 vectors at stride bn_ld: gamma, beta, running mean / var, count
  long long s_zc; int ld_zc;                                  // classifier input [z | 1]
  long long s_xn[NMB_MAX_HEAD];                               // normalised pre-activations (ld_hh)
  long long s_gate[NMB_MAX_HEAD];                             // relu gate x dropout scale (ld_hh)
  long long s_bn[NMB_MAX_HEAD];                               // [2][bn_ld]: batch inverse std, scratch
  long long s_dev;                                            // [B][4]: deviation health, disease, contrastive sign, label
  int drop_w;                                                 // sum of the hidden widths (row length of drop_keep)
  // NMB_FAMILY_DMVAE: S private + Zc shared latent dimensions
  int family, S, Zc, weighted;
  float beta;
  long long s_dzs;                                            // [B][Z]: shared part of d(total)/dz summed over the decoders
  long long scratch_floats;
};

// Device-side member record.
struct MemberDev {
  int arch_idx;
  int n_rows, batch;
  const float* xc[NMB_MAX_MOD];
  float* params; float* adam_m; float* adam_v; float* grads;
  const float* lr_steps;
  const float* y;           // head targets [n_rows]
  const int* row_order;     // [epochs][M][n_rows] or NULL
  const float* drop_keep;   // injected dropout keep flags or NULL
  long long n_drop_steps;
  unsigned long long seed;
  float lr, beta1, beta2, adam_eps;
  long long steps_done;     // mutable: minibatch steps taken so far
  int last_rows;            // mutable: rows of the last executed minibatch
  int last_slot;            // mutable: scratch slot that ran the last step
  long long launch_base;    // mutable: steps_done when the current pipelined launch started (work items wait on it)
};

// ---- layout (host) -------------------------------------------------------------------
inline int build_arch(const NmbArch& a, ArchDesc* d, const char** err) {
  auto fail = [&](const char* m) { *err = m; return 1; };
  if (a.n_mod < 1 || a.n_mod > NMB_MAX_MOD) return fail("n_mod out of range");
  if (a.n_hidden < 1 || a.n_hidden > NMB_MAX_HIDDEN) return fail("n_hidden out of range (1..4)");
  if (a.latent < 1 || a.c_dim < 0) return fail("bad latent / c_dim");
  if (a.combine < 0 || a.combine > NMB_COMBINE_MOPOE) return fail("No such combination method");
  if (a.loss_kind < 0 || a.loss_kind > NMB_LOSS_NEG_MSE) return fail("bad loss_kind");
  if (a.head_kind < NMB_HEAD_NONE || a.head_kind > NMB_HEAD_ENDTOEND) return fail("bad head_kind");
  if (a.family < NMB_FAMILY_CVAE || a.family > NMB_FAMILY_MVTCAE) return fail("bad family");
  if (a.family == NMB_FAMILY_MVTCAE && (a.head_kind != NMB_HEAD_NONE || a.loss_kind != NMB_LOSS_GAUSS_LL))
    return fail("NMB_FAMILY_MVTCAE: Gaussian log-likelihood, no head");
  if (a.family == NMB_FAMILY_DMVAE && (a.n_hidden != 2 || a.c_dim != 0 || a.head_kind != NMB_HEAD_NONE || a.s_dim < 0))
    return fail("NMB_FAMILY_DMVAE needs two hidden layers, c_dim == 0 (no covariates), no head and s_dim >= 0");
  if (a.head_kind == NMB_HEAD_ENDTOEND && 2 * a.n_mod > NMB_MAX_MOD) return fail("NMB_HEAD_ENDTOEND: at most 8 modalities");
  if (a.head_kind == NMB_HEAD_ENDTOEND && !(a.head_params[NMB_HP_DROPOUT] >= 0.f && a.head_params[NMB_HP_DROPOUT] < 1.f))
    return fail("dropout rate must be in [0, 1)");
  if (a.head_kind && (a.n_head_hidden < 1 || a.n_head_hidden > NMB_MAX_HEAD)) return fail("n_head_hidden out of range (1..3)");
  *d = ArchDesc{};
  d->M = a.n_mod; d->L = a.n_hidden; d->Z = a.latent; d->C = a.c_dim;
  d->combine = a.combine; d->loss_kind = a.loss_kind; d->non_linear = a.non_linear;
  for (int l = 0; l < a.n_hidden; ++l) {
    if (a.hidden[l] < 1) return fail("bad hidden width");
    d->hidden[l] = a.hidden[l];
  }
  long long off = 0, so = 0;
  const int B = kMaxBatch, L = d->L, Z = d->Z, C = d->C;
  auto lin = [&](int in, int out) {
    LinDesc r; r.in = in; r.out = out; r.ld = round4(in + 1); r.off = off;
    off += (long long)out * r.ld;
    return r;
  };
  auto buf = [&](int ld) { long long r = so; so += (long long)B * ld; return r; };
  int maxw = 2 * Z;
  d->MD = a.head_kind == NMB_HEAD_ENDTOEND ? 2 * d->M : d->M;
  for (int m = 0; m < d->MD; ++m) {
    ModDesc& q = d->mod[m];
    q.D = a.input_dims[m % d->M];
    if (q.D < 1) return fail("bad input dim");
    q.ldx = round4(q.D + C + 1);
    if (m < d->M) {
      for (int l = 0; l < L; ++l) q.enc[l] = lin(l == 0 ? q.D + C : d->hidden[l - 1], d->hidden[l]);
      q.head = lin(d->hidden[L - 1], 2 * Z);
    }
    for (int l = 0; l < L; ++l) q.dec[l] = lin(l == 0 ? Z + C : d->hidden[L - l], d->hidden[L - 1 - l]);
    q.outl = lin(d->hidden[0], q.D);
    q.lam_off = off; off += round4(q.D);
    for (int l = 0; l < L; ++l) {
      q.ld_h[l] = round4(d->hidden[l] + 1);          q.s_h[l] = buf(q.ld_h[l]);
      q.ld_k[l] = round4(d->hidden[L - 1 - l] + 1);  q.s_k[l] = buf(q.ld_k[l]);
      if (d->hidden[l] > maxw) maxw = d->hidden[l];
    }
    q.ld_mulv = round4(2 * Z); q.s_mulv = buf(q.ld_mulv); q.s_dmulv = buf(q.ld_mulv);
    q.ld_g0 = round4(Z + C + 1); q.s_g0 = buf(q.ld_g0);
    q.ld_xh = round4(q.D); q.s_xh = buf(q.ld_xh); q.s_xr = buf(q.ld_xh);
  }
  d->alpha_off = off; off += round4(d->M);
  d->head_kind = a.head_kind; d->HL = 0; d->sumD = 0; d->head_weight = a.head_weight;
  if (a.head_kind == NMB_HEAD_ENDTOEND) {
    d->HL = a.n_head_hidden;
    for (int i = 0; i < 6; ++i) d->hp[i] = a.head_params[i];
    d->ld_zc = round4(Z + 1); d->s_zc = buf(d->ld_zc);
    for (int l = 0; l <= d->HL; ++l) {
      if (l < d->HL) {
        if (a.head_hidden[l] < 1) return fail("bad head width");
        d->head_w[l] = a.head_hidden[l];
        d->drop_w += d->head_w[l];
        if (d->head_w[l] > maxw) maxw = d->head_w[l];
      }
      d->hd[l] = lin(l == 0 ? Z : d->head_w[l - 1], l == d->HL ? 2 : d->head_w[l]);
      if (l < d->HL) {
        d->bn_ld[l] = round4(d->head_w[l]); d->bn_off[l] = off; off += 5LL * d->bn_ld[l];
        d->ld_hh[l] = round4(d->head_w[l] + 1);
        d->s_hh[l] = buf(d->ld_hh[l]); d->s_xn[l] = buf(d->ld_hh[l]); d->s_gate[l] = buf(d->ld_hh[l]);
        d->s_bn[l] = so; so += 2LL * d->bn_ld[l];
      }
    }
    d->s_pred = buf(4); d->s_dpred = buf(4); d->s_dev = buf(4);
  } else if (a.head_kind) {     // appended after alpha: the packed layout of a head-less model is a prefix of this one
    d->HL = a.n_head_hidden;
    for (int m = 0; m < d->M; ++m) { d->mod[m].r_off = d->sumD; d->sumD += d->mod[m].D; d->mod[m].s_in = buf(d->mod[m].ldx); }
    for (int l = 0; l <= d->HL; ++l) {
      if (l < d->HL) {
        if (a.head_hidden[l] < 1) return fail("bad head width");
        d->head_w[l] = a.head_hidden[l];
        if (d->head_w[l] > maxw) maxw = d->head_w[l];
      }
      d->hd[l] = lin(l == 0 ? d->sumD : d->head_w[l - 1], l == d->HL ? 1 : d->head_w[l]);
    }
    d->ld_R = round4(d->sumD + 1); d->s_R = buf(d->ld_R); d->s_dR = buf(d->ld_R);
    for (int l = 0; l < d->HL; ++l) { d->ld_hh[l] = round4(d->head_w[l] + 1); d->s_hh[l] = buf(d->ld_hh[l]); }
    d->s_pred = buf(4); d->s_dpred = buf(4);
  }
  d->family = a.family; d->S = 0; d->Zc = Z; d->weighted = 0; d->beta = 1.f;
  if (a.family == NMB_FAMILY_DMVAE) {
    d->S = a.s_dim < Z ? a.s_dim : Z; d->Zc = Z - d->S; d->weighted = a.weighted ? 1 : 0; d->beta = a.beta;
    d->combine = NMB_COMBINE_POE; d->loss_kind = NMB_LOSS_NEG_MSE; d->non_linear = 2;      // ReLU everywhere
    d->s_dzs = buf(Z);
  }
  if (a.family == NMB_FAMILY_MVTCAE) d->beta = a.beta;
  d->n_params = off;
  d->s_mub = buf(Z); d->s_lvb = buf(Z); d->s_eps = buf(Z); d->s_dz = buf(Z);
  d->ld_g = round4(maxw);
  d->s_ga = buf(d->ld_g); d->s_gb = buf(d->ld_g);
  d->scratch_floats = (so + 3) & ~3LL;
  return 0;
}

// ---- Philox4x32-10 + Box-Muller (stream definition: oracle/philox.py) --------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// eps values of elements 4*group .. 4*group+3 of minibatch `step`.
__device__ inline void philox_normal4(unsigned long long seed, unsigned long long step,
                                      uint32_t stream, uint32_t group, float n[4]) {
  uint32_t w[4];
  philox4x32_10(group, (uint32_t)step, (uint32_t)(step >> 32), stream, (uint32_t)seed,
                (uint32_t)(seed >> 32), w);
  float u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = ((float)(w[i] >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
#pragma unroll
  for (int i = 0; i < 4; i += 2) {
    float r = sqrtf(-2.0f * logf(u[i]));
    float s, c;
    sincosf(6.283185307179586f * u[i + 1], &s, &c);
    n[i] = r * c; n[i + 1] = r * s;
  }
}

// ---- deterministic block reductions (fixed shuffle tree, fixed cross-warp order) -----------
__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// red must hold >= kThreads/32 + 1 floats.  All threads of the CTA must call.
__device__ inline float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    red[kThreads / 32] = s;
  }
  __syncthreads();
  return red[kThreads / 32];
}

}  // namespace nmb
