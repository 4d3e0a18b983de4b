"""Drop-in ``cVAE`` / ``cVAE_multimodal`` nn.Modules with the reference's constructor, forward and
loss signatures (reference ``cVAE.py``: Encoder 140-172, Decoder 174-206, cVAE 391-562,
cVAE_multimodal 1087-1211), backed by libnmb's fused CUDA kernels.

* Constructors draw from the torch RNG in exactly the reference's order -- including the
  discarded ``nn.Linear`` layers (cVAE.py:155-157, 190-191) and the Discriminator of ``cVAE``
  (cVAE.py:410) -- so ``torch.manual_seed(s)`` gives bit-identical initial weights and leaves the
  generator in the same state (eps draws line up too).  state_dict names match the reference.
* ``forward`` / ``forward_multimodal`` run ONE fused kernel that computes the forward pass, the
  loss and every gradient (nmb_ensemble_train with NO_ADAM | WRITE_GRADS); the returned loss
  tensors are autograd nodes whose backward hands the precomputed gradients to the parameters,
  so the reference loop  ``fwd -> loss -> zero_grad -> loss['total'].backward() -> optimizer1.step()``
  (multimodal_kfold_train_cvae_supervised.py:193-199) works unchanged.
* ``optimizer1`` is a ``torch.optim.Adam`` whose ``step()`` is the fused nmb_adam_step kernel.
* There is no eager / CPU fallback: on a CPU device these modules raise.

For throughput use ``EnsembleTrainer`` (ensemble.py), which runs whole training runs of many
such modules in a single launch; ``cVAE_multimodal.from_ensemble`` materialises its members back
into these classes (for ``torch.save`` / the test script).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn, optim
from torch.distributions import Normal

from . import _lib
from .ensemble import EnsembleTrainer, MemberSpec, pack_rows, _stream_ptr

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def compute_ll(x, x_recon):
    """cVAE.py:14-15 (host helper; the fused kernel computes the same quantity on the GPU)."""
    return x_recon.log_prob(x).sum(1, keepdims=True).mean(0)


class Encoder(nn.Module):
    """Parameter container with the reference's names and RNG order (cVAE.py:140-159)."""

    def __init__(self, input_dim, hidden_dim, c_dim, non_linear=False):
        super().__init__()
        self.input_size, self.hidden_dims, self.z_dim = input_dim, hidden_dim, hidden_dim[-1]
        self.c_dim, self.non_linear = c_dim, non_linear
        sizes = [input_dim + c_dim] + list(hidden_dim)
        self.layer_sizes_encoder = sizes
        layers = [nn.Linear(a, b, bias=True) for a, b in zip(sizes[:-1], sizes[1:])]
        self.encoder_layers = nn.Sequential(*layers[:-1])           # the last one is dropped
        self.enc_mean_layer = nn.Linear(sizes[-2], sizes[-1], bias=True)
        self.enc_logvar_layer = nn.Linear(sizes[-2], sizes[-1], bias=True)


class Decoder(nn.Module):
    """cVAE.py:174-194."""

    def __init__(self, input_dim, hidden_dim, c_dim, non_linear=False, init_logvar=-3):
        super().__init__()
        self.input_size, self.hidden_dims = input_dim, list(hidden_dim)[::-1]
        self.non_linear, self.init_logvar, self.c_dim = non_linear, init_logvar, c_dim
        sizes = self.hidden_dims + [input_dim]
        sizes[0] = self.hidden_dims[0] + c_dim
        self.layer_sizes_decoder = sizes
        layers = [nn.Linear(a, b, bias=True) for a, b in zip(sizes[:-1], sizes[1:])]
        self.decoder_layers = nn.Sequential(*layers[:-1])
        self.decoder_mean_layer = nn.Linear(sizes[-2], sizes[-1], bias=True)
        self.logvar_out = nn.Parameter(torch.FloatTensor(1, input_dim).fill_(init_logvar), requires_grad=True)


class Discriminator(nn.Module):
    """cVAE.py:210-227.  Never used by any script, but its construction consumes RNG inside
    ``cVAE.__init__`` (cVAE.py:410), so it is instantiated for seed-exact initialisation."""

    def __init__(self, input_dim, hidden_dim, c_dim, non_linear=False, init_logvar=-3):
        super().__init__()
        sizes = list(hidden_dim)[::-1] + [1]
        layers = [nn.Linear(a, b, bias=True) for a, b in zip(sizes[:-1], sizes[1:])]
        self.discriminator_layers = nn.Sequential(*layers[:-1])
        self.discriminator_mean_layer = nn.Linear(sizes[-2], sizes[-1], bias=True)


class _FusedStep(torch.autograd.Function):
    """Outputs of one fused forward+loss+backward launch as autograd nodes.  The kernel has already
    produced d(total)/d(param) for every parameter; backward scales them by d(loss)/d(total)."""

    @staticmethod
    def forward(ctx, owner, n_out, *params):
        outs, grads = owner._launch_step()
        ctx.grads = grads
        ctx.n_params = len(params)
        for o in outs[3:]:
            ctx.mark_non_differentiable(o)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        g_total, g_kl, g_ll = gouts[0], gouts[1], gouts[2]
        for g in (g_kl, g_ll):
            if g is not None and bool((g != 0).any()):
                raise RuntimeError("the fused cVAE step only supports backward through losses['total'] "
                                   "(the reference's training loop, train script :198)")
        scale = 1.0 if g_total is None else g_total.reshape(())
        return (None, None) + tuple(None if g is None else g * scale for g in ctx.grads)


class _FusedAdam(optim.Adam):
    """``torch.optim.Adam`` surface (param_groups, zero_grad, state) whose step() is nmb_adam_step."""

    def __init__(self, params, lr, owner):
        super().__init__(params, lr=lr)
        self._owner = owner

    @torch.no_grad()
    def step(self, closure=None):
        self._owner._adam_step(self.param_groups[0])


class _FusedBase(nn.Module):
    """Shared engine: a one-member EnsembleTrainer holding packed copies of the parameters."""

    _combine_default = "poe"
    _loss_kind = "gauss_ll"

    def _names(self):
        raise NotImplementedError

    def _trainable(self):
        """(name in packed layout, Parameter) in optimizer1 order."""
        raise NotImplementedError

    def _require_cuda(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("this cVAE runs only on CUDA (libnmb has no CPU fallback): call model.to('cuda')")
        return p.device

    def _engine_for(self, xs, cs, combine):
        dev = self._require_cuda()
        xc = [pack_rows(x.to(dev), c.to(dev)) for x, c in zip(xs, cs)]
        rows = xc[0].shape[0]
        if rows > 256:
            raise ValueError("a minibatch has at most 256 rows (train script :116)")
        key = (combine.lower(), rows, str(dev))
        eng = getattr(self, "_eng", None)
        if eng is None or self._eng_key != key:
            if eng is not None:
                eng.close()
            # the engine reads minibatches from its own persistent row buffers
            bufs = [torch.empty_like(t) for t in xc]
            spec = MemberSpec(self._dims, self._hidden, self.latent_dim, self.c_dim, bufs, combine=combine,
                              loss_kind=self._loss_kind, non_linear=self._non_linear, batch=rows,
                              lr=self.learning_rate)
            eng = EnsembleTrainer([spec], device=dev, keep_grads=True)
            object.__setattr__(self, "_eng", eng)
            object.__setattr__(self, "_eng_key", key)
            object.__setattr__(self, "_eng_xc", bufs)
        for dst, src in zip(self._eng_xc, xc):
            dst.copy_(src)
        eng.load_state_dict(0, self._packed_state())
        return eng

    def _packed_state(self):
        return {k: p.detach() for k, p in self._trainable_named()}

    def _launch_step(self):
        eng, eps = self._pending
        eng.grads.zero_()
        losses = eng.train_steps(1, eps=eps[None, None], record_losses=True,
                                 flags=_lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS)
        mu, lv, xr = eng.peek(0)
        g = eng.state_dict(0, "grads")
        grads = []
        for name, p in self._trainable_named():
            grads.append(g[name].reshape(p.shape))
        lo = losses[0, 0]
        return [lo[0].reshape(1), lo[1].reshape(()), lo[2].reshape(1), mu, lv] + list(xr), grads

    def _fused_forward(self, xs, cs, combine):
        eng = self._engine_for(xs, cs, combine)
        rows = xs[0].shape[0]
        # the reference draws eps with randn_like(mu) from the global generator (cVAE.py:1132)
        eps = torch.randn((rows, self.latent_dim), device=xs[0].device if xs[0].is_cuda else eng.device,
                          dtype=torch.float32)
        object.__setattr__(self, "_pending", (eng, eps.to(eng.device)))
        params = [p for _, p in self._trainable_named()]
        outs = _FusedStep.apply(self, 0, *params)
        return outs

    def _adam_step(self, group):
        eng = getattr(self, "_eng", None)
        if eng is None:
            raise RuntimeError("optimizer1.step() called before any forward pass")
        named = list(self._trainable_named())
        # gradients as the user left them (zero_grad / backward), packed into the engine layout
        eng.grads.zero_()
        gviews = eng._views(0, eng.grads)
        pviews = eng._views(0, eng.params)
        for name, p in named:
            pviews[name].copy_(p.detach().reshape(pviews[name].shape))
            if p.grad is not None:
                gviews[name].copy_(p.grad.reshape(gviews[name].shape))
        t = getattr(self, "_adam_t", 0) + 1
        object.__setattr__(self, "_adam_t", t)
        b1, b2 = group["betas"]
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.nmb_adam_step(eng.params.data_ptr(), eng.grads.data_ptr(), eng.adam_m.data_ptr(),
                                             eng.adam_v.data_ptr(), eng.total_params, t, float(group["lr"]),
                                             float(b1), float(b2), float(group["eps"]), _stream_ptr(eng.device)),
                       "nmb_adam_step")
        for name, p in named:
            p.copy_(pviews[name].reshape(p.shape))

    def _trainable_named(self):
        return self._trainable()

    def __getstate__(self):      # torch.save(model): engines are not picklable and are rebuilt lazily
        d = self.__dict__.copy()
        for k in ("_eng", "_eng_key", "_eng_xc", "_pending"):
            d.pop(k, None)
        return d


class cVAE(_FusedBase):
    """Single-modality conditional VAE (cVAE.py:391-562)."""

    def __init__(self, input_dim, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=4,
                 non_linear=False):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim + [latent_dim]
        self.latent_dim, self.c_dim, self.modalities, self.learning_rate = latent_dim, c_dim, modalities, learning_rate
        self._dims, self._hidden, self._non_linear = [input_dim], list(hidden_dim), bool(non_linear)
        self.encoder = Encoder(input_dim, self.hidden_dim, c_dim, non_linear)
        self.decoder = Decoder(input_dim, self.hidden_dim, c_dim, non_linear)
        self.discriminator = Discriminator(input_dim, self.hidden_dim, c_dim, non_linear)
        self.optimizer1 = _FusedAdam(list(self.encoder.parameters()) + list(self.decoder.parameters()),
                                     lr=learning_rate, owner=self)

    def _trainable(self):
        for k, p in self.encoder.named_parameters():
            yield "encoder_list.0." + k, p
        for k, p in self.decoder.named_parameters():
            yield "decoder_list.0." + k, p

    def _packed_state(self):
        sd = super()._packed_state()
        sd["alpha_m_list.0"] = torch.zeros(1)
        return sd

    def forward(self, x, c):
        self.zero_grad()
        total, kl, ll, mu, logvar, xr = self._fused_forward([x], [c], "poe")
        scale = self.decoder.logvar_out.detach().exp().pow(0.5)
        object.__setattr__(self, "_last_losses", {"total": total, "kl": kl, "ll": ll})
        return {"x_recon": Normal(loc=xr, scale=scale), "mu": mu, "logvar": logvar}

    def loss_function(self, x, fwd_rtn):
        """{'total','kl','ll'} of the forward pass that produced fwd_rtn (cVAE.py:491-504)."""
        return dict(self._last_losses)

    def calc_kl(self, mu, logvar):
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean(0)

    def calc_ll(self, x, x_recon):
        return compute_ll(x, x_recon)

    def _recon(self, x, c, mode):
        dev = self._require_cuda()
        xc = [pack_rows(torch.as_tensor(np.asarray(x), dtype=torch.float32).to(dev),
                        torch.as_tensor(np.asarray(c)).to(dev))]
        sd = self._packed_state()
        eng = EnsembleTrainer([MemberSpec(self._dims, self._hidden, self.latent_dim, self.c_dim, xc,
                                          non_linear=self._non_linear, state_dict=sd)], device=dev)
        out = eng.reconstruct([xc], mode=mode, want_latent=True)
        torch.cuda.synchronize(dev)
        eng.close()
        return out

    def pred_latent(self, x, c, DEVICE=None):
        """(mu, exp(logvar)) as numpy (cVAE.py:540-547)."""
        x = x.to_numpy() if hasattr(x, "to_numpy") else x
        _, mu, lv = self._recon(x, c, "mean")
        return mu[0].cpu().numpy(), lv[0].exp().cpu().numpy()

    def pred_recon(self, x, c, DEVICE=None):
        """Decode the latent MEAN (cVAE.py:549-555)."""
        x = x.to_numpy() if hasattr(x, "to_numpy") else x
        xhat, _, _ = self._recon(x, c, "mean")
        return xhat[0][0].cpu().numpy()


class cVAE_multimodal(_FusedBase):
    """M encoders + M decoders with latent fusion (cVAE.py:1087-1211)."""

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=3,
                 non_linear=False):
        super().__init__()
        self.input_dim_list = input_dim_list
        self.hidden_dim = hidden_dim + [latent_dim]
        self.latent_dim, self.c_dim, self.modalities, self.learning_rate = latent_dim, c_dim, modalities, learning_rate
        self._dims = [int(d) for d in input_dim_list[:modalities]]
        self._hidden, self._non_linear = list(hidden_dim), bool(non_linear)
        # RNG order of cVAE.py:1107-1109: alphas, encoders, decoders
        self.alpha_m_list = nn.ParameterList(
            [nn.Parameter(torch.randn(1, requires_grad=True)) for _ in range(modalities)])
        self.encoder_list = nn.ModuleList(
            [Encoder(input_dim_list[i], self.hidden_dim, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list = nn.ModuleList(
            [Decoder(input_dim_list[i], self.hidden_dim, c_dim, non_linear) for i in range(modalities)])
        self.optimizer1 = _FusedAdam(
            [p for m in self.encoder_list for p in m.parameters()]
            + [p for m in self.decoder_list for p in m.parameters()]
            + list(self.alpha_m_list.parameters()), lr=learning_rate, owner=self)

    def _trainable(self):
        for i, m in enumerate(self.encoder_list):
            for k, p in m.named_parameters():
                yield f"encoder_list.{i}.{k}", p
        for i, m in enumerate(self.decoder_list):
            for k, p in m.named_parameters():
                yield f"decoder_list.{i}.{k}", p
        for i, p in enumerate(self.alpha_m_list):
            yield f"alpha_m_list.{i}", p

    def forward_multimodal(self, xes, cs, combine):
        self.zero_grad()
        _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, combine)   # ValueError if unknown
        outs = self._fused_forward(list(xes), list(cs), combine)
        total, kl, ll, mu, logvar = outs[:5]
        x_recons = [Normal(loc=outs[5 + i], scale=self.decoder_list[i].logvar_out.detach().exp().pow(0.5))
                    for i in range(self.modalities)]
        object.__setattr__(self, "_last_losses", {"total": total, "kl": kl, "ll": ll})
        return {"x_recons": x_recons, "mu_multimodal": mu, "logvar_multimodal": logvar}

    def loss_function_multimodal(self, xes, fwd_rtn):
        """sum_m (kl - ll_m) of the forward pass that produced fwd_rtn (cVAE.py:1187-1196)."""
        return dict(self._last_losses)

    def calc_kl(self, mu, logvar):
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean(0)

    def calc_ll(self, x, x_recon):
        return compute_ll(x, x_recon)

    def pred_recon(self, xes, c, DEVICE=None, combine="poe"):
        """Test-time reconstruction; z is SAMPLED with torch.randn like the reference (cVAE.py:1198-1208)."""
        dev = self._require_cuda()
        c_t = torch.as_tensor(np.asarray(c)).to(dev)
        xc = [pack_rows(torch.as_tensor(np.asarray(getattr(x, "values", x)), dtype=torch.float32).to(dev), c_t)
              for x in xes[: self.modalities]]
        eng = EnsembleTrainer([MemberSpec(self._dims, self._hidden, self.latent_dim, self.c_dim, xc, combine=combine,
                                          non_linear=self._non_linear, state_dict=self._packed_state())], device=dev)
        eps = torch.randn((xc[0].shape[0], self.latent_dim), dtype=torch.float32)   # CPU generator, like :1207
        xhat, _, _ = eng.reconstruct([xc], mode="sample", eps=[eps.to(dev)])
        torch.cuda.synchronize(dev)
        eng.close()
        return [t.cpu().numpy() for t in xhat[0]]

    def reconstruction_deviation_multimodal(self, xes, x_preds):
        """Per-subject sum_d (x - xhat)^2 / D (cVAE.py:1210-1211), float64 like the reference's host math."""
        return [np.sum((np.asarray(xes[m], dtype=np.float64) - x_preds[m]) ** 2, axis=1) / np.asarray(xes[m]).shape[1]
                for m in range(self.modalities)]

    @classmethod
    def from_ensemble(cls, trainer: EnsembleTrainer, i: int):
        """Materialise ensemble member i as a reference-shaped module (for torch.save / pred_recon)."""
        s = trainer.specs[i]
        model = cls(list(s.input_dims), list(s.hidden), s.latent, s.c_dim, learning_rate=s.lr,
                    modalities=len(s.input_dims), non_linear=s.non_linear)
        sd = trainer.state_dict(i)
        model.load_state_dict({k: v.cpu() for k, v in sd.items()})
        return model.to(trainer.device)
