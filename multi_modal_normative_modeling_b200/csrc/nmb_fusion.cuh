// Device helpers shared by the fused training kernels: 16-wide row segments and the latent
// fusion operators of cVAE.py:1144-1164 (forward and hand-derived backward, oracle/cvae_numpy.py).
#pragma once
#include "nmb_common.cuh"

namespace nmb {

// ---- 16-wide row segments (tensor-core engine epilogues) -------------------------------------
// p is 16-byte aligned; only the first `nvalid` of the 16 floats exist / may be written.
__device__ __forceinline__ void load16(const float* p, int nvalid, float fill, float (&o)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (4 * q + 3 < nvalid) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
      o[4 * q] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[4 * q + j] = (4 * q + j < nvalid) ? p[4 * q + j] : fill;
    }
  }
}
__device__ __forceinline__ void store16(float* p, int nvalid, const float (&o)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (4 * q + 3 < nvalid) {
      *reinterpret_cast<float4*>(p + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (4 * q + j < nvalid) p[4 * q + j] = o[4 * q + j];
    }
  }
}


// ---- latent fusion (cVAE.py:1144-1164) -----------------------------------------------------
// Forward for one (row, z) element.  mu[m], lv[m] are the per-modality heads.
struct Fused { float mu, lv; };

// (.cg loads: alpha is updated every step, possibly by another SM between two work items of the member)
__device__ inline void softmax_alpha(const float* alpha, int M, float* out) {
  float mx = __ldcg(alpha);
  for (int m = 1; m < M; ++m) mx = fmaxf(mx, __ldcg(alpha + m));
  float s = 0.f;
  for (int m = 0; m < M; ++m) { out[m] = expf(__ldcg(alpha + m) - mx); s += out[m]; }
  for (int m = 0; m < M; ++m) out[m] /= s;
}

__device__ inline Fused fuse_forward(const float* mu, const float* lv, int M, int combine, const float* w) {
  Fused f;
  if (M == 1) { f.mu = mu[0]; f.lv = lv[0]; return f; }
  if (combine == NMB_COMBINE_MOE) {
    float sm = 0.f, sv = 0.f;
    for (int m = 0; m < M; ++m) { sm += mu[m]; sv += expf(lv[m]); }
    f.mu = sm / M; f.lv = logf(sv / M);
    return f;
  }
  float s = 0.f, num = 0.f, sm = 0.f, sv = 0.f;
  for (int m = 0; m < M; ++m) {
    const float v = expf(lv[m]);
    const float t = (combine == NMB_COMBINE_GPOE ? w[m] : 1.f) / v;
    s += t; num += mu[m] * t; sm += mu[m]; sv += v;
  }
  const float pmu = num / s, pvar = 1.f / s;
  if (combine == NMB_COMBINE_MOPOE) { f.mu = (sm + pmu) / (M + 1); f.lv = logf((sv + pvar) / (M + 1)); }
  else { f.mu = pmu; f.lv = logf(pvar); }
  return f;
}

// Backward for one element: given d_mu_bar, d_lv_bar produce d_mu[m], d_lv[m] and the per-element
// contribution to d(alpha_softmax_weight)[m] (gPoE).  See oracle/cvae_numpy.py:fuse_backward.
__device__ inline void fuse_backward(const float* mu, const float* lv, int M, int combine, const float* w,
                                     float dmu_bar, float dlv_bar, float* dmu, float* dlv, float* dw) {
  if (M == 1) { dmu[0] = dmu_bar; dlv[0] = dlv_bar; return; }
  float v[NMB_MAX_MOD];
  float sv = 0.f, sm = 0.f;
  for (int m = 0; m < M; ++m) { v[m] = expf(lv[m]); sv += v[m]; sm += mu[m]; }
  if (combine == NMB_COMBINE_MOE) {
    const float var_bar = sv / M;
    const float dvar = dlv_bar / var_bar;
    for (int m = 0; m < M; ++m) { dmu[m] = dmu_bar / M; dlv[m] = dvar / M * v[m]; }
    return;
  }
  float s = 0.f, num = 0.f;
  for (int m = 0; m < M; ++m) {
    const float t = (combine == NMB_COMBINE_GPOE ? w[m] : 1.f) / v[m];
    s += t; num += mu[m] * t;
  }
  const float pmu = num / s, pvar = 1.f / s;
  float dpm, dpv, base_mu = 0.f, base_v = 0.f;
  if (combine == NMB_COMBINE_MOPOE) {
    const float var_bar = (sv + pvar) / (M + 1);
    const float dvar = dlv_bar / var_bar;
    base_mu = dmu_bar / (M + 1); base_v = dvar / (M + 1);
    dpm = base_mu; dpv = base_v;
  } else {
    dpm = dmu_bar; dpv = dlv_bar / pvar;
  }
  for (int m = 0; m < M; ++m) {
    const float a = (combine == NMB_COMBINE_GPOE ? w[m] : 1.f);
    const float t = a / v[m];
    const float dt = dpm * (mu[m] - pmu) / s - dpv * pvar * pvar;
    dmu[m] = base_mu + dpm * t / s;
    const float dv = base_v - dt * a / (v[m] * v[m]);
    dlv[m] = dv * v[m];
    if (dw) dw[m] = dt / v[m];
  }
}


// ---- mvtCAE's fusion as written (cVAE.py:1778, 1795-1816) ---------------------------------------------------------
// 'poe' hands VARIANCES v_m = exp(lv_m) to ProductOfExperts2, which treats them as log-variances: T_m = 1 / exp(v_m),
// mu = sum mu_m T_m / sum T_m, "variance" = log(1 / sum T_m).  Every branch then clamps the variance to >= 1e-6.
// Returns the fused (mu, log clamp(var)); *clamped tells the backward pass that the variance path carries no gradient.
__device__ inline Fused fuse_forward_mvtcae(const float* mu, const float* lv, int M, int combine, const float* w, bool* clamped) {
  Fused f;
  float var;
  if (combine == NMB_COMBINE_POE) {
    float s = 0.f, num = 0.f;
    for (int m = 0; m < M; ++m) { const float t = expf(-expf(lv[m])); s += t; num += mu[m] * t; }
    f.mu = num / s;
    var = logf(1.f / s);
  } else {
    const int M1 = M;      // (mvtCAE has no single-modality shortcut)
    if (M1 == 1 && combine != NMB_COMBINE_MOPOE) { f.mu = mu[0]; var = expf(lv[0]); }
    else { const Fused g = fuse_forward(mu, lv, M, combine, w); f.mu = g.mu; var = expf(g.lv); }
  }
  *clamped = !(var >= 1e-6f);
  f.lv = logf(*clamped ? 1e-6f : var);
  return f;
}

__device__ inline void fuse_backward_mvtcae(const float* mu, const float* lv, int M, int combine, const float* w, bool clamped,
                                            float dmu_bar, float dlv_bar, float* dmu, float* dlv, float* dw) {
  if (clamped) dlv_bar = 0.f;
  if (combine != NMB_COMBINE_POE) {
    if (dw) for (int m = 0; m < M; ++m) dw[m] = 0.f;
    fuse_backward(mu, lv, M, combine, w, dmu_bar, dlv_bar, dmu, dlv, dw);
    return;
  }
  float T[NMB_MAX_MOD], v[NMB_MAX_MOD], s = 0.f, num = 0.f;
  for (int m = 0; m < M; ++m) { v[m] = expf(lv[m]); T[m] = expf(-v[m]); s += T[m]; num += mu[m] * T[m]; }
  const float pmu = num / s, var = logf(1.f / s);
  const float ds = clamped ? 0.f : -(dlv_bar / var) / s;         // var = -log s
  for (int m = 0; m < M; ++m) {
    dmu[m] = dmu_bar * T[m] / s;
    const float dT = dmu_bar * (mu[m] - pmu) / s + ds;
    dlv[m] = dT * (-T[m]) * v[m];
    if (dw) dw[m] = 0.f;
  }
}

}  // namespace nmb
