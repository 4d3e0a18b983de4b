"""Eager-PyTorch restatement of the reference cVAE path (CPU).  TEST INFRASTRUCTURE ONLY.

This is the "port" CPU baseline and the per-step parity checker.  It restates, in
compact form, the behaviour of the reference classes:

* ``Encoder``            cVAE.py:140-172
* ``Decoder``            cVAE.py:174-206
* ``cVAE``               cVAE.py:391-443, 491-504, 549-555
* fusion ops             cVAE.py:986-1083, gPoE inline 1154-1157
* ``cVAE_multimodal``    cVAE.py:1087-1211
* nmmlp -MSE variant     multimodal_kfold_cvae_nmmlp.py:124-127

Parameter names, shapes and -- critically -- the torch-RNG draw order of the
constructors (including the *discarded* ``nn.Linear`` layers of cVAE.py:155-157,
190-191, 220-227) are preserved so that ``torch.manual_seed(s)`` followed by
construction yields bit-identical weights and leaves the generator in the same
state as the reference (verified by tests/test_oracle_vs_reference_golden.py
against vectors produced by oracle/make_golden.py from the real reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

LOG_2PI = math.log(2.0 * math.pi)


def _burn_linear(n_in: int, n_out: int) -> None:
    """Consume exactly the RNG draws of one ``nn.Linear(n_in, n_out)`` (weight, bias)."""
    nn.Linear(n_in, n_out, bias=True)


class OracleEncoder(nn.Module):
    """cVAE.py:140-172.  ``hidden_dim`` already includes the latent size as its last entry."""

    def __init__(self, input_dim, hidden_dim, c_dim, non_linear=False):
        super().__init__()
        sizes = [input_dim + c_dim] + list(hidden_dim)
        body = [nn.Linear(a, b) for a, b in zip(sizes[:-2], sizes[1:-1])]
        _burn_linear(sizes[-2], sizes[-1])            # the discarded last layer (cVAE.py:155-157)
        self.encoder_layers = nn.Sequential(*body)
        self.enc_mean_layer = nn.Linear(sizes[-2], sizes[-1])
        self.enc_logvar_layer = nn.Linear(sizes[-2], sizes[-1])
        self.non_linear = non_linear
        self.c_dim = c_dim

    def forward(self, x, c):
        h = torch.cat((x, c), dim=1)                  # int64 c promotes to float32 (cVAE.py:163)
        for layer in self.encoder_layers:
            h = layer(h)
            if self.non_linear:
                h = F.leaky_relu(h)                   # slope 0.01
        return self.enc_mean_layer(h), self.enc_logvar_layer(h)


class OracleDecoder(nn.Module):
    """cVAE.py:174-206.  Returns (mu_out, logvar_out) instead of a Normal object."""

    def __init__(self, input_dim, hidden_dim, c_dim, non_linear=False, init_logvar=-3.0):
        super().__init__()
        rev = list(hidden_dim)[::-1]
        sizes = rev + [input_dim]
        sizes[0] = rev[0] + c_dim
        body = [nn.Linear(a, b) for a, b in zip(sizes[:-2], sizes[1:-1])]
        _burn_linear(sizes[-2], sizes[-1])            # discarded (cVAE.py:190-191)
        self.decoder_layers = nn.Sequential(*body)
        self.decoder_mean_layer = nn.Linear(sizes[-2], sizes[-1])
        self.logvar_out = nn.Parameter(torch.full((1, input_dim), float(init_logvar)))
        self.non_linear = non_linear
        self.c_dim = c_dim

    def forward(self, z, c):
        g = torch.cat((z, c.reshape(-1, self.c_dim)), dim=1)
        for layer in self.decoder_layers:
            g = layer(g)
            if self.non_linear:
                g = F.leaky_relu(g)
        return self.decoder_mean_layer(g), self.logvar_out


def _burn_discriminator(hidden_dim) -> None:
    """RNG draws of ``Discriminator.__init__`` (cVAE.py:210-227): sizes rev(hidden)+[1]."""
    sizes = list(hidden_dim)[::-1] + [1]
    for a, b in zip(sizes[:-1], sizes[1:]):
        _burn_linear(a, b)
    _burn_linear(sizes[-2], sizes[-1])


def gauss_ll(x, mu_out, logvar_out):
    """``compute_ll`` cVAE.py:14-15 on ``Normal(mu_out, exp(logvar_out)**0.5)``; shape [1]."""
    var = logvar_out.exp()
    lp = -((x - mu_out) ** 2) / (2.0 * var) - 0.5 * logvar_out - 0.5 * LOG_2PI
    return lp.sum(1, keepdim=True).mean(0)


def neg_mse_ll(x, mu_out):
    """nmmlp variant: ``-MSELoss(mean)`` (multimodal_kfold_cvae_nmmlp.py:124-127); scalar."""
    return -((mu_out - x) ** 2).mean()


def kl_term(mu, logvar):
    """cVAE.py:429-430 / 1138-1139."""
    return (-0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1)).mean(0)


def fuse_latent(mus, variances, combine, alphas=None):
    """``combine_latent`` cVAE.py:1144-1164 with the expert ops of cVAE.py:986-1083.

    mus, variances: [M, B, Z].  alphas: list of M tensors of shape [1] (gPoE only).
    """
    if mus.shape[0] == 1:
        return mus[0], variances[0]
    kind = combine.lower()
    m = mus.shape[0]
    if kind == "poe":
        t = 1.0 / variances
        return (mus * t).sum(0) / t.sum(0), 1.0 / t.sum(0)
    if kind == "gpoe":
        a = torch.softmax(torch.stack(list(alphas)), dim=0).reshape(m, 1, 1)
        w = a / variances
        return (mus * a / variances).sum(0) / w.sum(0), 1.0 / w.sum(0)
    if kind == "moe":
        return mus.sum(0) / m, variances.sum(0) / m
    if kind == "mopoe":
        t = 1.0 / variances
        p_mu, p_var = (mus * t).sum(0) / t.sum(0), 1.0 / t.sum(0)
        return (mus.sum(0) + p_mu) / (m + 1), (variances.sum(0) + p_var) / (m + 1)
    raise ValueError("No such combination method")


class OracleCVAE(nn.Module):
    """Single-modality ``cVAE`` (cVAE.py:391-443).  Discriminator draws are burned, not kept."""

    def __init__(self, input_dim, hidden_dim, latent_dim, c_dim, learning_rate=1e-4,
                 modalities=4, non_linear=False):
        super().__init__()
        hd = list(hidden_dim) + [latent_dim]
        self.encoder = OracleEncoder(input_dim, hd, c_dim, non_linear)
        self.decoder = OracleDecoder(input_dim, hd, c_dim, non_linear)
        _burn_discriminator(hd)
        self.optimizer1 = torch.optim.Adam(
            list(self.encoder.parameters()) + list(self.decoder.parameters()), lr=learning_rate)

    def step_losses(self, x, c, eps=None):
        """forward (cVAE.py:435-443) + loss_function (cVAE.py:491-504)."""
        mu, logvar = self.encoder(x, c)
        if eps is None:
            eps = torch.randn_like(mu)
        z = mu + eps * torch.exp(0.5 * logvar)
        mu_out, lv_out = self.decoder(z, c)
        kl = kl_term(mu, logvar)
        ll = gauss_ll(x, mu_out, lv_out)
        return {"total": kl - ll, "kl": kl, "ll": ll, "mu": mu, "logvar": logvar, "x_recon": mu_out}

    def pred_recon(self, x, c):
        """cVAE.py:549-555: decode the *mean*."""
        with torch.no_grad():
            mu, _ = self.encoder(x, c)
            return self.decoder(mu, c)[0]


class OracleCVAEMultimodal(nn.Module):
    """``cVAE_multimodal`` (cVAE.py:1087-1211).  loss_kind: 'gauss_ll' | 'neg_mse'."""

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=1e-4,
                 modalities=3, non_linear=False, loss_kind="gauss_ll", rng_order="cvae"):
        super().__init__()
        hd = list(hidden_dim) + [latent_dim]
        self.modalities = modalities
        self.loss_kind = loss_kind
        # RNG order: alphas, then encoders, then decoders (cVAE.py:1107-1109).  The class inside
        # multimodal_kfold_cvae_nmmlp.py (:57-83) draws encoders, decoders, alphas and then its (unused) MLP.
        if rng_order == "cvae":
            self.alpha_m_list = nn.ParameterList(
                [nn.Parameter(torch.randn(1)) for _ in range(modalities)])
        self.encoder_list = nn.ModuleList(
            [OracleEncoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list = nn.ModuleList(
            [OracleDecoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        if rng_order != "cvae":
            self.alpha_m_list = nn.ParameterList(
                [nn.Parameter(torch.randn(1)) for _ in range(modalities)])
            _burn_linear(sum(input_dim_list[:modalities]), 128)
            _burn_linear(128, 64)
            _burn_linear(64, 1)
        self.optimizer1 = torch.optim.Adam(
            [p for e in self.encoder_list for p in e.parameters()]
            + [p for d in self.decoder_list for p in d.parameters()]
            + list(self.alpha_m_list.parameters()), lr=learning_rate)

    def latent(self, xs, cs, combine):
        enc = [self.encoder_list[i](xs[i], cs[i]) for i in range(self.modalities)]
        mus = torch.stack([e[0] for e in enc])
        variances = torch.exp(torch.stack([e[1] for e in enc]))
        mu_mm, var_mm = fuse_latent(mus, variances, combine, list(self.alpha_m_list))
        return mu_mm, torch.log(var_mm)

    def step_losses(self, xs, cs, combine, eps=None):
        """forward_multimodal (cVAE.py:1166-1182) + loss_function_multimodal (1187-1196)."""
        mu_mm, logvar_mm = self.latent(xs, cs, combine)
        if eps is None:
            eps = torch.randn_like(mu_mm)
        z = mu_mm + eps * torch.exp(0.5 * logvar_mm)
        recons = [self.decoder_list[i](z, cs[i]) for i in range(self.modalities)]
        total = kl_sum = ll_sum = 0
        for i in range(self.modalities):
            kl = kl_term(mu_mm, logvar_mm)
            if self.loss_kind == "gauss_ll":
                ll = gauss_ll(xs[i], recons[i][0], recons[i][1])
            else:
                ll = neg_mse_ll(xs[i], recons[i][0])
            total = total + (kl - ll)
            kl_sum = kl_sum + kl
            ll_sum = ll_sum + ll
        return {"total": total, "kl": kl_sum, "ll": ll_sum, "mu": mu_mm, "logvar": logvar_mm,
                "x_recons": [r[0] for r in recons]}

    def pred_recon(self, xs, c, combine, eps=None):
        """cVAE.py:1198-1208: z is *sampled* at test time (eps injectable for determinism)."""
        with torch.no_grad():
            cs = [c] * self.modalities
            mu_mm, logvar_mm = self.latent(xs, cs, combine)
            if eps is None:
                eps = torch.randn_like(mu_mm)
            z = mu_mm + eps * torch.exp(0.5 * logvar_mm)
            return [self.decoder_list[i](z, c)[0] for i in range(self.modalities)]


def reference_train_loop(model, xs, cs, combine, epochs, batch_size=256, eps_fn=None,
                         lr_fn=None):
    """The hot loop of multimodal_kfold_train_cvae_supervised.py:177-199 (no shuffle, last
    batch partial, LR untouched because ``optimizer1.lr = clr`` is a no-op, :183) and, when
    ``lr_fn`` is given, of multimodal_kfold_cvae_nmmlp.py:374-400 (LR set per step through
    param_groups).  xs / cs: lists of [N, D_m] / [N, C] tensors.  Returns per-step losses.
    """
    n = xs[0].shape[0]
    log = []
    step = 0
    for _ in range(epochs):
        for lo in range(0, n, batch_size):
            step += 1
            if lr_fn is not None:
                for g in model.optimizer1.param_groups:
                    g["lr"] = lr_fn(step)
            xb = [x[lo:lo + batch_size] for x in xs]
            cb = [c[lo:lo + batch_size] for c in cs]
            eps = None if eps_fn is None else eps_fn(step - 1, xb[0].shape[0])
            out = model.step_losses(xb, cb, combine, eps)
            model.optimizer1.zero_grad()
            out["total"].backward()
            model.optimizer1.step()
            log.append((float(out["total"]), float(out["kl"]), float(out["ll"])))
    return np.asarray(log, dtype=np.float64)


def cyclic_lr(step, n_samples, batch_size=256, base_lr=1e-6, max_lr=5e-5, gamma=0.98):
    """Triangular cyclic LR of multimodal_kfold_cvae_nmmlp.py:363-381 (1-based global step)."""
    step_size = 2 * np.ceil(n_samples / batch_size)
    cycle = np.floor(1 + step / (2 * step_size))
    x_lr = np.abs(step / step_size - 2 * cycle + 1)
    return float(base_lr + (max_lr - base_lr) * max(0, 1 - x_lr) * gamma ** cycle)


class OracleCVAERegression(OracleCVAEMultimodal):
    """``cVAE_multimodal_regression`` (cVAE.py:2211-2347): the multimodal cVAE plus an MLP on the concatenated
    residuals x_m - x_recon_m.loc that predicts a scalar target (FI).  RNG order of the constructor: encoders,
    decoders, alphas, regressor (:2231-2255); Adam over encoders, decoders, regressor, alphas (:2260-2266)."""

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=1e-4, modalities=3, non_linear=False,
                 head_hidden=(128, 64)):
        nn.Module.__init__(self)
        hd = list(hidden_dim) + [latent_dim]
        self.modalities = modalities
        self.loss_kind = "gauss_ll"
        self.encoder_list = nn.ModuleList(
            [OracleEncoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list = nn.ModuleList(
            [OracleDecoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        self.alpha_m_list = nn.ParameterList([nn.Parameter(torch.randn(1)) for _ in range(modalities)])
        widths = [sum(input_dim_list[:modalities])] + list(head_hidden)
        layers = []
        for a, b in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        layers.append(nn.Linear(widths[-1], 1))
        self.regressor = nn.Sequential(*layers)
        self.optimizer1 = torch.optim.Adam(
            list(self.encoder_list.parameters()) + list(self.decoder_list.parameters())
            + list(self.regressor.parameters()) + list(self.alpha_m_list.parameters()), lr=learning_rate)

    def step_losses(self, xs, cs, combine, eps=None, true_fi=None, lambda_reg=1.0):
        """forward_multimodal (:2309-2332) + loss_function_multimodal (:2334-2347)."""
        out = super().step_losses(xs, cs, combine, eps)
        resid = torch.cat([xs[i] - out["x_recons"][i] for i in range(self.modalities)], dim=1)
        out["fi_pred"] = self.regressor(resid)
        if true_fi is not None:
            out["regression"] = torch.mean((out["fi_pred"].squeeze() - true_fi.squeeze()) ** 2)      # nn.MSELoss
            out["total"] = out["total"] + lambda_reg * out["regression"]
        return out


def regression_train_loop(model, xs, cs, fi, order, combine, batch_size, eps_steps):
    """Loop body of multimodal_kfold_train_cvae_supervised_regression.py:119-131 with the shuffling loaders made
    explicit: order[epoch][m] is the permutation modality m's DataLoader(shuffle=True) yields in that epoch (one
    independent permutation per modality, :94); the target comes from modality 0's loader (:125).
    Returns per-step (total, kl, ll, regression)."""
    n = xs[0].shape[0]
    log, step = [], 0
    for ep in range(order.shape[0]):
        for lo in range(0, n, batch_size):
            idx = [torch.as_tensor(order[ep, m, lo:lo + batch_size], dtype=torch.long) for m in range(model.modalities)]
            xb = [xs[m][idx[m]] for m in range(model.modalities)]
            cb = [cs[idx[m]] for m in range(model.modalities)]
            out = model.step_losses(xb, cb, combine, torch.as_tensor(eps_steps[step][: xb[0].shape[0]]), fi[idx[0]])
            model.optimizer1.zero_grad()
            out["total"].backward()
            model.optimizer1.step()
            log.append((float(out["total"]), float(out["kl"]), float(out["ll"]), float(out["regression"])))
            step += 1
    return np.asarray(log, dtype=np.float64)


class OracleCVAEEndToEnd(nn.Module):
    """``cVAE_multimodal_endtoend`` v2 (cVAE.py:2021-2207) + ``Classifier`` (:2004-2018): shared encoders, plain PoE
    fusion, a health and a disease decoder set, a classifier on z; loss = w_rec (rec_h + rec_d) + w_kl kl + CE +
    w_con * hinge(margin +- (dev_h - dev_d)).  RNG order of the constructor: encoders, health decoders, disease decoders,
    classifier (:2042-2052)."""

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=1e-4, modalities=3, non_linear=False,
                 classifier_layers=(128, 64), dropout_rate=0.5):
        super().__init__()
        hd = list(hidden_dim) + [latent_dim]
        self.modalities, self.dropout_rate = modalities, dropout_rate
        self.encoder_list = nn.ModuleList(
            [OracleEncoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list_health = nn.ModuleList(
            [OracleDecoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list_disease = nn.ModuleList(
            [OracleDecoder(input_dim_list[i], hd, c_dim, non_linear) for i in range(modalities)])
        sizes = [latent_dim] + list(classifier_layers)
        layers = []
        for a, b in zip(sizes[:-1], sizes[1:]):
            layers += [nn.Linear(a, b), nn.BatchNorm1d(b), nn.ReLU(), nn.Dropout(dropout_rate)]
        layers.append(nn.Linear(sizes[-1], 2))
        self.classifier = nn.Module()
        self.classifier.classifier = nn.Sequential(*layers)
        self.optimizer = torch.optim.Adam(self.parameters(), lr=learning_rate)

    def classify(self, z, keep=None):
        """keep: list of [rows, width] keep flags, one per Dropout (training mode); None = nn.Dropout's own draws."""
        h, k = z, 0
        for mod in self.classifier.classifier:
            if isinstance(mod, nn.Dropout) and keep is not None and self.training and mod.p > 0:
                h = h * keep[k] / (1.0 - mod.p)
                k += 1
            else:
                h = mod(h)
        return h

    def latent(self, xs, cs):
        enc = [self.encoder_list[i](xs[i], cs[i]) for i in range(self.modalities)]
        mus, logvars = torch.stack([e[0] for e in enc]), torch.stack([e[1] for e in enc])
        t = 1 / torch.exp(logvars)
        return torch.sum(mus * t, dim=0) / torch.sum(t, dim=0), torch.log(1 / torch.sum(t, dim=0))      # :2081-2088

    def step_losses(self, xs, cs, labels=None, eps=None, keep=None, margin=1.0, w_con=0.1, w_kl=0.1, w_rec=0.1):
        mu, logvar = self.latent(xs, cs)
        if eps is None:
            eps = torch.randn_like(mu)
        z = mu + eps * torch.exp(0.5 * logvar)
        rh = [self.decoder_list_health[i](z, cs[i]) for i in range(self.modalities)]
        rd = [self.decoder_list_disease[i](z, cs[i]) for i in range(self.modalities)]
        out = {"mu": mu, "logvar": logvar, "logits": self.classify(z, keep),
               "x_recons_health": [r[0] for r in rh], "x_recons_disease": [r[0] for r in rd]}
        if labels is None:
            return out
        rec_h = sum(-gauss_ll(xs[i], rh[i][0], rh[i][1]) for i in range(self.modalities))
        rec_d = sum(-gauss_ll(xs[i], rd[i][0], rd[i][1]) for i in range(self.modalities))
        dev_h = torch.stack([((xs[i] - rh[i][0]) ** 2).mean(dim=1) for i in range(self.modalities)]).mean(dim=0)
        dev_d = torch.stack([((xs[i] - rd[i][0]) ** 2).mean(dim=1) for i in range(self.modalities)]).mean(dim=0)
        lab = labels.to(mu.dtype)
        con = torch.mean((1 - lab) * torch.relu(margin + dev_h - dev_d) + lab * torch.relu(margin + dev_d - dev_h))
        kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean()
        ce = torch.nn.functional.cross_entropy(out["logits"], labels)
        out.update({"total": w_rec * (rec_h + rec_d) + w_kl * kl + ce + w_con * con, "kl": kl, "ce": ce, "rec_health": rec_h,
                    "rec_disease": rec_d, "contrastive": con})
        return out

    def predict(self, xs, cs):
        """:2198-2203: classifier(mu_combined), to be called in eval mode."""
        with torch.no_grad():
            return self.classify(self.latent(xs, cs)[0])


def e2e_train_loop(model, xs, cs, labels, batch_size, epochs, eps_steps, keep_steps, widths, margin, w_con):
    """Loop body of multimodal_kfold_cvae_nmpmcont.py:226-247 (shuffle=False, last batch partial; the cyclic learning rate
    written to ``optimizer.lr`` never reaches Adam).  Returns per-step (total, kl, ce, rec_health, rec_disease, contrastive)."""
    n = xs[0].shape[0]
    log, step = [], 0
    for _ in range(epochs):
        for lo in range(0, n, batch_size):
            xb = [x[lo:lo + batch_size] for x in xs]
            cb = [cs[lo:lo + batch_size]] * model.modalities
            rows = xb[0].shape[0]
            keep, o = [], 0
            for w in widths:
                keep.append(torch.as_tensor(keep_steps[step][:rows, o:o + w])); o += w
            out = model.step_losses(xb, cb, labels[lo:lo + batch_size], torch.as_tensor(eps_steps[step][:rows]), keep, margin, w_con)
            model.optimizer.zero_grad()
            out["total"].backward()
            model.optimizer.step()
            log.append([float(out[k].detach()) for k in ("total", "kl", "ce", "rec_health", "rec_disease", "contrastive")])
            step += 1
    return np.asarray(log, dtype=np.float64)


class OracleDMVAE(nn.Module):
    """The DMVAE family of the baseline zoo: ``DMVAE`` (cVAE.py:1491-1618), ``mmVAEPlus`` (:1895-2002; identical code with
    beta = 0.05) and ``WeightedDMVAE`` (:1620-1752; ``weights = |randn(M)|`` multiply each modality's kl and ll, beta unused).
    VariationalEncoder / VariationalDecoder (:1454-1480): two ReLU layers, no covariates, sigmoid output.  The first s_dim
    latent dimensions are private (passed on as their mean), the rest shared (ProductOfExperts2, :1482-1489)."""

    def __init__(self, input_dim_list, hidden_dim, latent_dim, s_dim, learning_rate=1e-4, modalities=3, beta=1.0, weighted=False):
        super().__init__()
        self.modalities, self.s_dim, self.beta, self.weighted = modalities, s_dim, beta, weighted

        def enc(d):
            m = nn.Module()
            m.fc1, m.fc2 = nn.Linear(d, hidden_dim[0]), nn.Linear(hidden_dim[0], hidden_dim[1])
            m.fc_mu, m.fc_logvar = nn.Linear(hidden_dim[1], latent_dim), nn.Linear(hidden_dim[1], latent_dim)
            return m

        def dec(d):
            m = nn.Module()
            m.fc1, m.fc2, m.fc_out = nn.Linear(latent_dim, hidden_dim[1]), nn.Linear(hidden_dim[1], hidden_dim[0]), nn.Linear(hidden_dim[0], d)
            return m
        self.encoder_list = nn.ModuleList([enc(input_dim_list[i]) for i in range(modalities)])
        self.decoder_list = nn.ModuleList([dec(input_dim_list[i]) for i in range(modalities)])
        if weighted:
            self.weights = nn.Parameter(torch.abs(torch.randn(modalities)))
        self.optimizer1 = torch.optim.Adam(self.parameters(), lr=learning_rate)

    def step_losses(self, xs, eps=None):
        mu_s, mu_c, lv_c = [], [], []
        for i, e in enumerate(self.encoder_list):
            h = torch.relu(e.fc2(torch.relu(e.fc1(xs[i]))))
            mu, lv = e.fc_mu(h), e.fc_logvar(h)
            mu_s.append(mu[:, : self.s_dim]); mu_c.append(mu[:, self.s_dim:]); lv_c.append(lv[:, self.s_dim:])
        mu_c, lv_c = torch.stack(mu_c), torch.stack(lv_c)
        t = 1.0 / torch.exp(lv_c)
        mu = torch.sum(mu_c * t, dim=0) / torch.sum(t, dim=0)
        logvar = torch.log(1.0 / torch.sum(t, dim=0))
        if eps is None:
            eps = torch.randn_like(mu)
        z = mu + eps * torch.exp(0.5 * logvar)
        recons = []
        for i, d in enumerate(self.decoder_list):
            h = torch.relu(d.fc2(torch.relu(d.fc1(torch.cat((z, mu_s[i]), dim=1)))))
            recons.append(torch.sigmoid(d.fc_out(h)))
        kl = ll = 0
        for i in range(self.modalities):
            w = self.weights[i] if self.weighted else 1.0
            kl = kl + w * (-0.5 * torch.sum(1 + logvar - mu.pow(2) - torch.exp(logvar), dim=1).mean(0))
            ll = ll + w * (-0.5 * torch.sum((xs[i] - recons[i]) ** 2, dim=1).mean(0))
        total = (kl - ll) if self.weighted else (kl * self.beta - ll)
        return {"total": total, "kl": kl, "ll": ll, "mu_c": mu, "logvar_c": logvar, "x_recons": recons}


def dmvae_train_loop(model, xs, batch_size, epochs, eps_steps, zc):
    """Loop body of the train script (:177-199) for a DMVAE-family model.  Returns per-step (total, kl, ll)."""
    n = xs[0].shape[0]
    log, step = [], 0
    for _ in range(epochs):
        for lo in range(0, n, batch_size):
            xb = [x[lo:lo + batch_size] for x in xs]
            out = model.step_losses(xb, torch.as_tensor(eps_steps[step][: xb[0].shape[0], :zc]))
            model.optimizer1.zero_grad()
            out["total"].backward()
            model.optimizer1.step()
            log.append((float(out["total"].detach()), float(torch.as_tensor(out["kl"]).detach()), float(out["ll"].detach())))
            step += 1
    return np.asarray(log, dtype=np.float64)


class OracleMvtCAE(OracleCVAEMultimodal):
    """``mvtCAE`` (cVAE.py:1754-1893) as written: the cVAE_multimodal architecture; the fused variance is clamped to >= 1e-6
    (:1815); the 'poe' branch hands VARIANCES to ProductOfExperts2, which exponentiates them again and returns a
    log-variance that is then used as the variance (:1778, :1800); total = sum_m (kl + 1e-5 ll_m + beta tc) with the
    log-likelihood entering with a plus sign (:1862) and tc = - sum_i mean_m logsumexp_b mu_m[b, i] (:1846-1853)."""

    beta = 0.0001

    def latent(self, xs, cs, combine):
        enc = [self.encoder_list[i](xs[i], cs[i]) for i in range(self.modalities)]
        mus = torch.stack([e[0] for e in enc])
        variances = torch.exp(torch.stack([e[1] for e in enc]))
        if combine.lower() == "poe":
            t = 1.0 / torch.exp(variances)
            mu_mm, var_mm = torch.sum(mus * t, dim=0) / torch.sum(t, dim=0), torch.log(1.0 / torch.sum(t, dim=0))
        else:
            mu_mm, var_mm = fuse_latent(mus, variances, combine, list(self.alpha_m_list))
        self._mus = mus
        return mu_mm, torch.log(torch.clamp(var_mm, min=1e-6))

    def step_losses(self, xs, cs, combine, eps=None):
        out = super().step_losses(xs, cs, combine, eps)      # its kl / ll sums are the reference's losses['kl'] / ['ll']
        m = self.modalities
        tc = 0
        for i in range(out["mu"].shape[1]):
            tc = tc - torch.stack([self._mus[j][:, i].logsumexp(dim=0) for j in range(m)]).mean(dim=0)
        out["tc"] = m * tc
        out["total"] = out["kl"] + 0.00001 * out["ll"] + self.beta * out["tc"]
        return out
