"""Hand-derived forward / backward / Adam of the cVAE path in numpy.  TEST INFRASTRUCTURE ONLY.

The CUDA kernels implement exactly these formulas; this file is the
readable statement of them, checked against autograd of the reference
(tests/test_oracle_math.py, tests/test_oracle_vs_reference_golden.py).

Reference behaviour restated:
* Encoder.forward / Decoder.forward         cVAE.py:161-172, 197-206
* reparameterise                            cVAE.py:1130-1133
* calc_kl / compute_ll                      cVAE.py:1138-1139, 14-15
* combine_latent + expert ops               cVAE.py:1144-1164, 986-1083
* loss_function_multimodal                  cVAE.py:1187-1196
* -MSE recon (nmmlp)                        multimodal_kfold_cvae_nmmlp.py:124-127
* torch.optim.Adam defaults (lr, betas .9/.999, eps 1e-8; cVAE.py:1111-1116)

Parameters are handled as a dict with the reference's state_dict names
(``encoder_list.{m}.encoder_layers.{l}.weight`` ...), values numpy arrays.
"""
from __future__ import annotations

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)
SLOPE = 0.01  # F.leaky_relu default negative slope


def lrelu(a, on):
    return np.where(a > 0, a, SLOPE * a) if on else a


def lrelu_grad(a, on):
    return np.where(a > 0, 1.0, SLOPE).astype(a.dtype) if on else np.ones_like(a)


def _n_hidden(params, prefix):
    n = 0
    while f"{prefix}.{n}.weight" in params:
        n += 1
    return n


def encoder_forward(p, m, x, c, non_linear):
    """Returns (mu, logvar, cache) for modality m."""
    pre = f"encoder_list.{m}"
    h = np.concatenate([x, c.astype(x.dtype)], axis=1)
    acts, pres = [h], []
    for l in range(_n_hidden(p, f"{pre}.encoder_layers")):
        a = acts[-1] @ p[f"{pre}.encoder_layers.{l}.weight"].T + p[f"{pre}.encoder_layers.{l}.bias"]
        pres.append(a)
        acts.append(lrelu(a, non_linear))
    mu = acts[-1] @ p[f"{pre}.enc_mean_layer.weight"].T + p[f"{pre}.enc_mean_layer.bias"]
    lv = acts[-1] @ p[f"{pre}.enc_logvar_layer.weight"].T + p[f"{pre}.enc_logvar_layer.bias"]
    return mu, lv, (acts, pres)


def decoder_forward(p, m, z, c, non_linear):
    pre = f"decoder_list.{m}"
    g = np.concatenate([z, c.astype(z.dtype)], axis=1)
    acts, pres = [g], []
    for l in range(_n_hidden(p, f"{pre}.decoder_layers")):
        a = acts[-1] @ p[f"{pre}.decoder_layers.{l}.weight"].T + p[f"{pre}.decoder_layers.{l}.bias"]
        pres.append(a)
        acts.append(lrelu(a, non_linear))
    xh = acts[-1] @ p[f"{pre}.decoder_mean_layer.weight"].T + p[f"{pre}.decoder_mean_layer.bias"]
    return xh, (acts, pres)


def fuse_forward(mus, lvs, combine, alphas):
    """mus, lvs: [M,B,Z].  Returns (mu_bar, logvar_bar, cache)."""
    m = mus.shape[0]
    if m == 1:
        return mus[0], lvs[0], None
    v = np.exp(lvs)
    kind = combine.lower()
    if kind in ("poe", "gpoe", "mopoe"):
        if kind == "gpoe":
            a = np.exp(alphas - alphas.max())
            a = (a / a.sum()).reshape(m, 1, 1)
        else:
            a = np.ones((m, 1, 1), dtype=mus.dtype)
        t = a / v
        s = t.sum(0)
        p_mu, p_var = (mus * t).sum(0) / s, 1.0 / s
        if kind == "mopoe":
            mu_bar = (mus.sum(0) + p_mu) / (m + 1)
            var_bar = (v.sum(0) + p_var) / (m + 1)
        else:
            mu_bar, var_bar = p_mu, p_var
        cache = (v, a, t, s, p_mu, p_var)
    elif kind == "moe":
        mu_bar, var_bar = mus.sum(0) / m, v.sum(0) / m
        cache = (v,)
    else:
        raise ValueError("No such combination method")
    return mu_bar, np.log(var_bar), (cache, var_bar)


def fuse_backward(d_mu_bar, d_lv_bar, mus, lvs, combine, alphas, fcache):
    """Returns (d_mus [M,B,Z], d_lvs [M,B,Z], d_alphas [M])."""
    m = mus.shape[0]
    d_alpha = np.zeros(m, dtype=mus.dtype)
    if m == 1:
        return d_mu_bar[None], d_lv_bar[None], d_alpha
    cache, var_bar = fcache
    d_var_bar = d_lv_bar / var_bar
    kind = combine.lower()
    if kind == "moe":
        (v,) = cache
        d_mus = np.broadcast_to(d_mu_bar / m, mus.shape).copy()
        d_v = np.broadcast_to(d_var_bar / m, mus.shape).copy()
        return d_mus, d_v * v, d_alpha
    v, a, t, s, p_mu, p_var = cache
    if kind == "mopoe":
        d_mus = np.broadcast_to(d_mu_bar / (m + 1), mus.shape).copy()
        d_v = np.broadcast_to(d_var_bar / (m + 1), mus.shape).copy()
        d_pmu, d_pvar = d_mu_bar / (m + 1), d_var_bar / (m + 1)
    else:
        d_mus = np.zeros_like(mus)
        d_v = np.zeros_like(mus)
        d_pmu, d_pvar = d_mu_bar, d_var_bar
    # product expert: p_mu = sum(mu t)/s, p_var = 1/s, t = a / v
    d_mus += d_pmu * t / s
    d_t = d_pmu * (mus - p_mu) / s - d_pvar * p_var * p_var
    d_v += -d_t * a / (v * v)
    if kind == "gpoe":
        d_a = (d_t / v).sum(axis=(1, 2))
        a1 = a.reshape(m)
        d_alpha = a1 * (d_a - (a1 * d_a).sum())
    return d_mus, d_v * v, d_alpha


def step(p, xs, cs, eps, combine="poe", non_linear=True, loss_kind="gauss_ll"):
    """One forward + loss + backward.  xs, cs: lists over modalities; eps: [B,Z].

    Returns (losses dict, outputs dict, grads dict keyed like p).
    """
    m_n = len(xs)
    b = xs[0].shape[0]
    dt = xs[0].dtype
    enc = [encoder_forward(p, m, xs[m], cs[m], non_linear) for m in range(m_n)]
    mus = np.stack([e[0] for e in enc])
    lvs = np.stack([e[1] for e in enc])
    alphas = (np.array([p[f"alpha_m_list.{m}"][0] for m in range(m_n)], dtype=dt)
              if "alpha_m_list.0" in p else np.zeros(m_n, dtype=dt))
    mu_bar, lv_bar, fcache = fuse_forward(mus, lvs, combine, alphas)
    s = np.exp(0.5 * lv_bar)
    z = mu_bar + eps * s
    dec = [decoder_forward(p, m, z, cs[m], non_linear) for m in range(m_n)]

    kl = (-0.5 * np.sum(1 + lv_bar - mu_bar ** 2 - np.exp(lv_bar), axis=1)).mean()
    grads = {}
    ll_sum = 0.0
    d_z = np.zeros_like(z)
    for m in range(m_n):
        xh, (acts, pres) = dec[m]
        pre = f"decoder_list.{m}"
        lam = p[f"{pre}.logvar_out"]
        r = xs[m] - xh
        if loss_kind == "gauss_ll":
            var = np.exp(lam)
            ll = (-(r * r) / (2 * var) - 0.5 * lam - 0.5 * LOG_2PI).sum(1).mean()
            d_xh = -(r / var) / b                         # d(total)/d xh
            grads[f"{pre}.logvar_out"] = (0.5 * (1 - r * r / var)).sum(0, keepdims=True) / b
        else:
            d = xs[m].shape[1]
            ll = -(r * r).mean()
            d_xh = -2.0 * r / (b * d)
            grads[f"{pre}.logvar_out"] = np.zeros_like(lam)
        ll_sum += ll
        grads[f"{pre}.decoder_mean_layer.weight"] = d_xh.T @ acts[-1]
        grads[f"{pre}.decoder_mean_layer.bias"] = d_xh.sum(0)
        d_act = d_xh @ p[f"{pre}.decoder_mean_layer.weight"]
        for l in reversed(range(len(pres))):
            d_pre = d_act * lrelu_grad(pres[l], non_linear)
            grads[f"{pre}.decoder_layers.{l}.weight"] = d_pre.T @ acts[l]
            grads[f"{pre}.decoder_layers.{l}.bias"] = d_pre.sum(0)
            d_act = d_pre @ p[f"{pre}.decoder_layers.{l}.weight"]
        d_z += d_act[:, : z.shape[1]]

    # total = sum_m (kl - ll_m)  ->  kl is counted M times (cVAE.py:1189-1195)
    d_mu_bar = d_z + m_n * mu_bar / b
    d_lv_bar = d_z * eps * s * 0.5 + m_n * (np.exp(lv_bar) - 1) / (2 * b)
    d_mus, d_lvs, d_alpha = fuse_backward(d_mu_bar, d_lv_bar, mus, lvs, combine, alphas, fcache)
    for m in range(m_n):
        if f"alpha_m_list.{m}" in p:
            grads[f"alpha_m_list.{m}"] = np.array([d_alpha[m]], dtype=dt)
        _, _, (acts, pres) = enc[m]
        pre = f"encoder_list.{m}"
        grads[f"{pre}.enc_mean_layer.weight"] = d_mus[m].T @ acts[-1]
        grads[f"{pre}.enc_mean_layer.bias"] = d_mus[m].sum(0)
        grads[f"{pre}.enc_logvar_layer.weight"] = d_lvs[m].T @ acts[-1]
        grads[f"{pre}.enc_logvar_layer.bias"] = d_lvs[m].sum(0)
        d_act = d_mus[m] @ p[f"{pre}.enc_mean_layer.weight"] + d_lvs[m] @ p[f"{pre}.enc_logvar_layer.weight"]
        for l in reversed(range(len(pres))):
            d_pre = d_act * lrelu_grad(pres[l], non_linear)
            grads[f"{pre}.encoder_layers.{l}.weight"] = d_pre.T @ acts[l]
            grads[f"{pre}.encoder_layers.{l}.bias"] = d_pre.sum(0)
            d_act = d_pre @ p[f"{pre}.encoder_layers.{l}.weight"]

    losses = {"total": m_n * kl - ll_sum, "kl": m_n * kl, "ll": ll_sum}
    outs = {"mu": mu_bar, "logvar": lv_bar, "z": z, "x_recons": [d[0] for d in dec],
            "mus": mus, "logvars": lvs}
    return losses, outs, grads


def adam_update(p, g, m, v, t, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update, step count t (1-based), in place.

    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
    """
    m *= beta1
    m += (1 - beta1) * g
    v *= beta2
    v += (1 - beta2) * g * g
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p -= (lr / bc1) * (m / denom)


def train(p, xs, cs, eps_steps, combine, epochs, batch=256, lr=1e-4, lr_steps=None,
          non_linear=True, loss_kind="gauss_ll", skip_no_grad=("alpha",)):
    """The reference hot loop (train script :177-199) on numpy arrays.

    eps_steps: [n_steps, batch, Z].  Parameters whose gradient is None in the reference
    (alpha_m when M == 1, cVAE.py:1146-1147) are skipped by Adam.
    Returns per-step (total, kl, ll).
    """
    n = xs[0].shape[0]
    m_state = {k: np.zeros_like(a) for k, a in p.items()}
    v_state = {k: np.zeros_like(a) for k, a in p.items()}
    log, t = [], 0
    uses_alpha = len(xs) > 1 and combine.lower() == "gpoe"
    for _ in range(epochs):
        for lo in range(0, n, batch):
            xb = [x[lo:lo + batch] for x in xs]
            cb = [c[lo:lo + batch] for c in cs]
            eps = eps_steps[t][: xb[0].shape[0]]
            losses, _, g = step(p, xb, cb, eps, combine, non_linear, loss_kind)
            t += 1
            cur_lr = lr if lr_steps is None else lr_steps[t - 1]
            for k in p:
                if k.startswith("alpha_m_list") and not uses_alpha:
                    continue
                if loss_kind != "gauss_ll" and k.endswith("logvar_out"):
                    continue
                adam_update(p[k], g[k], m_state[k], v_state[k], t, cur_lr)
            log.append((losses["total"], losses["kl"], losses["ll"]))
    return np.asarray(log)
