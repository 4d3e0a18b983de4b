"""Drop-ins for the DMVAE family of the reference's baseline zoo (``-Model DMVAE | mmVAEPlus | WeightedDMVAE``
in the train script :141-148): ``VariationalEncoder`` / ``VariationalDecoder`` / ``ProductOfExperts2`` (cVAE.py:1454-1489),
``DMVAE`` (:1491-1618), ``WeightedDMVAE`` (:1620-1752) and ``mmVAEPlus`` (:1895-2002).

As written in the reference these models take no covariates (the ``c`` arguments are ignored), split every modality's
latent into ``s_dim = c_dim`` private dimensions -- handed to the decoder as their MEAN -- and ``latent_dim - c_dim``
shared ones (product of experts, reparameterised), and decode through a sigmoid.  With the train script's c_dim = 29 and
latent_dim = 10 the shared part is empty: M deterministic autoencoders, kl = 0.  The fused step (``NMB_FAMILY_DMVAE``)
covers both situations; ``mmJSD`` lives in ``.cVAE`` (it is the PoE cVAE)."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch
from torch import nn

from . import _lib
from .cVAE import _FusedAdam, _FusedBase, _as_float_cuda
from .ensemble import EnsembleTrainer, MemberSpec, pack_rows


class VariationalEncoder(nn.Module):
    """Parameter container (cVAE.py:1454-1467); evaluated inside the fused kernels."""

    def __init__(self, input_dim, hidden_dims, latent_dim, s_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dims[0])
        self.fc2 = nn.Linear(hidden_dims[0], hidden_dims[1])
        self.fc_mu = nn.Linear(hidden_dims[1], latent_dim)
        self.fc_logvar = nn.Linear(hidden_dims[1], latent_dim)


class VariationalDecoder(nn.Module):
    """Parameter container (cVAE.py:1469-1480)."""

    def __init__(self, output_dim, hidden_dims, combined_dim):
        super().__init__()
        self.fc1 = nn.Linear(combined_dim, hidden_dims[1])
        self.fc2 = nn.Linear(hidden_dims[1], hidden_dims[0])
        self.fc_out = nn.Linear(hidden_dims[0], output_dim)


class ProductOfExperts2(nn.Module):
    """(mu, logvar) of the product of the experts N(mu[m], exp(logvar[m])) (cVAE.py:1482-1489)."""

    def forward(self, mu, logvar):
        var_inv = 1.0 / torch.exp(logvar)
        return torch.sum(mu * var_inv, dim=0) / torch.sum(var_inv, dim=0), torch.log(1.0 / torch.sum(var_inv, dim=0))


class DMVAE(_FusedBase):
    """cVAE.py:1491-1618.  ``forward_multimodal(xes, cs, combine)`` (cs and combine are ignored, like the reference) runs the
    fused forward + loss + backward launch; ``loss_function_multimodal(xes, fwd_rtn)`` returns its losses;
    ``losses['total'].backward()`` hands out the gradients; ``optimizer1.step()`` is the fused Adam."""

    _beta = 1.0
    _weighted = False

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=3, non_linear=False):
        super().__init__()
        if len(hidden_dim) < 2:
            raise ValueError("VariationalEncoder / VariationalDecoder use hidden_dim[0] and hidden_dim[1]")
        self.input_dim_list, self.hidden_dim, self.latent_dim = input_dim_list, hidden_dim, latent_dim
        self.c_dim = self.s_dim = c_dim
        self.beta = self._beta
        self.modalities, self.learning_rate, self.non_linear = modalities, learning_rate, non_linear
        self._dims = [int(d) for d in input_dim_list[:modalities]]
        self._hidden = [int(hidden_dim[0]), int(hidden_dim[1])]
        self.encoder_list = nn.ModuleList([VariationalEncoder(input_dim_list[i], hidden_dim, latent_dim, c_dim)
                                           for i in range(modalities)])
        self.decoder_list = nn.ModuleList([VariationalDecoder(input_dim_list[i], hidden_dim, latent_dim)
                                           for i in range(modalities)])
        self.join_z = ProductOfExperts2()
        self._extra_init()
        self.optimizer1 = _FusedAdam(list(self.parameters()), lr=learning_rate, owner=self)

    def _extra_init(self):
        pass

    # ---- layout ------------------------------------------------------------------------------------------------
    def _trainable(self):          # optim.Adam(self.parameters()) order
        return list(self.named_parameters())

    def _family_kwargs(self):
        return dict(family="dmvae", s_dim=int(self.s_dim), weighted=bool(self._weighted), beta=float(self.beta))

    def _shared(self):
        return max(0, int(self.latent_dim) - int(self.s_dim))

    def _make_engine(self, dev, dims, combine, rows, keep_grads, names=None, with_head=False):
        bufs = [torch.zeros((rows, _lib.packed_row_stride(int(d), 0)), dtype=torch.float32, device=dev) for d in dims]
        spec = MemberSpec(dims, self._hidden, int(self.latent_dim), 0, bufs, batch=rows, lr=self.learning_rate,
                          **self._family_kwargs())
        eng = EnsembleTrainer([spec], device=dev, keep_grads=keep_grads)
        eng._rows_buf = bufs
        return eng

    def _moments(self, dev):
        if self.__dict__.get("_adam_m") is None:
            n = _lib.arch_param_count(_lib.make_arch(self._dims, self._hidden, int(self.latent_dim), 0, **self._family_kwargs()))
            object.__setattr__(self, "_adam_m", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_v", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_t", 0)
            return self._adam_m, self._adam_v
        return super()._moments(dev)

    def _engine_for(self, xs, kind="train"):
        dev = self._require_cuda()
        rows = int(xs[0].shape[0])
        if kind == "train" and not 1 <= rows <= 256:
            raise ValueError("a minibatch has 1..256 rows (train script :116)")
        key = (kind, rows if kind == "train" else 0, str(dev)) + tuple(sorted(self._family_kwargs().items()))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, self._dims, "poe", rows if kind == "train" else 1,
                                                         keep_grads=kind == "train")
        none = torch.zeros((rows, 0), device=dev)
        packed = [pack_rows(_as_float_cuda(x, dev), none, out=dst if kind == "train" else None)
                  for x, dst in zip(xs, eng._rows_buf)]
        self._load_weights(eng)
        return eng, packed

    # ---- the training-loop surface -----------------------------------------------------------------------------------
    def _launch_step(self):
        eng, eps = self._pending
        eng.grads.zero_()
        losses = eng.train_steps(1, eps=eps[None, None], record_losses=True,
                                 flags=_lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS | int(self._engine_flags))
        mu, lv, xr = eng.peek(0)
        rows, zc = xr[0].shape[0], self._shared()
        mu_c, lv_c = mu.reshape(-1)[: rows * zc].view(rows, zc), lv.reshape(-1)[: rows * zc].view(rows, zc)   # stored [rows][Zc]
        gviews = eng.__dict__.setdefault("_gviews", eng._views(0, eng.grads))
        grads = [gviews[name].view(p.shape) for name, p in self._trainable_named()]
        lo = losses[0, 0]
        return [lo[0].reshape(()), lo[1].reshape(()), lo[2].reshape(()), mu_c, lv_c] + list(xr), grads

    def forward_multimodal(self, xes, cs=None, combine=None):
        from .cVAE import _FusedStep
        xs = list(xes)[: self.modalities]
        eng, _ = self._engine_for(xs)
        rows, zc = xs[0].shape[0], self._shared()
        eps = torch.zeros((rows, int(self.latent_dim)), dtype=torch.float32, device=eng.device)
        eps[:, :zc] = torch.randn((rows, zc), device=eng.device, dtype=torch.float32)        # randn_like(mu_c) (:1531-1533)
        object.__setattr__(self, "_pending", (eng, eps))
        outs = _FusedStep.apply(self, 0, *[p for _, p in self._trainable_named()])
        fwd = {"x_recons": list(outs[5:]), "mu_c": outs[3], "logvar_c": outs[4]}
        self._remember(fwd, {"total": outs[0], "kl": outs[1], "ll": outs[2]}, "mu_c")
        return fwd

    def loss_function_multimodal(self, xes, fwd_rtn):
        return self._losses_of(fwd_rtn, "mu_c")

    def encode(self, x, c, m):
        raise NotImplementedError("the DMVAE family is evaluated inside the fused step: use forward_multimodal / pred_recon")

    def reparameterize(self, mu, logvar):
        return mu + torch.randn_like(mu) * torch.exp(0.5 * logvar)

    def pred_recon(self, xes, cs=None, device=None, combine=None):
        """Reconstructions of every modality (numpy), shared z sampled like the reference (:1574-1596)."""
        xs = [x.values if isinstance(x, pd.DataFrame) else x for x in list(xes)[: self.modalities]]
        eng, packed = self._engine_for(xs, kind="infer")
        rows, zc = packed[0].shape[0], self._shared()
        eps = torch.zeros((rows, int(self.latent_dim)), dtype=torch.float32, device=eng.device)
        eps[:, :zc] = torch.randn((rows, zc), device=eng.device, dtype=torch.float32)
        xhat, _, _ = eng.reconstruct([packed], mode="sample", eps=[eps], engine="fp32" if int(self._engine_flags) & _lib.TRAIN_FP32 else "tcs")
        torch.cuda.synchronize(eng.device)
        return [t.cpu().numpy() for t in xhat[0]]

    def configure_optimizers(self):
        return self.optimizer1

    @classmethod
    def from_ensemble(cls, trainer: EnsembleTrainer, i: int):
        """Materialise ensemble member i as a reference-shaped module (for torch.save / pred_recon)."""
        s = trainer.specs[i]
        model = cls(list(s.input_dims), list(s.hidden), s.latent, s.s_dim, learning_rate=s.lr, modalities=len(s.input_dims),
                    non_linear=True)
        model.load_state_dict({k: v.cpu() for k, v in trainer.state_dict(i).items()}, strict=True)
        return model.to(trainer.device)

    def reconstruction_deviation_multimodal(self, xes, x_preds):
        out = []
        for m in range(self.modalities):
            x = xes[m].values if isinstance(xes[m], pd.DataFrame) else np.asarray(xes[m])
            p = x_preds[m].values if isinstance(x_preds[m], pd.DataFrame) else np.asarray(x_preds[m])
            out.append(np.sum((x - p) ** 2, axis=1) / x.shape[1])
        return out


class mmVAEPlus(DMVAE):
    """cVAE.py:1895-2002: the DMVAE code with beta = 0.05."""
    _beta = 0.05


class WeightedDMVAE(DMVAE):
    """cVAE.py:1620-1752: learnable positive-initialised weights multiply each modality's kl and ll (total = kl - ll)."""
    _weighted = True

    def _extra_init(self):
        self.weights = nn.Parameter(torch.abs(torch.randn(self.modalities)))
