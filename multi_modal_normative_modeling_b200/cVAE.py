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
        if g_total is None or bool((g_total.reshape(()) == 1).item()):
            return (None, None) + tuple(ctx.grads)           # the views of the packed gradient buffer themselves
        scale = g_total.reshape(())
        return (None, None) + tuple(None if g is None else g * scale for g in ctx.grads)


class _FusedAdam(optim.Adam):
    """``torch.optim.Adam`` surface (param_groups, zero_grad, state, state_dict) whose step() is the
    nmb_adam_step kernel.  exp_avg / exp_avg_sq live in module-owned packed buffers (they survive changes of the
    minibatch size, ``.to(device)`` and ``torch.save``); ``state[p]`` exposes them as views."""

    def __init__(self, params, lr, owner):
        super().__init__(params, lr=lr)
        self._owner = owner

    @torch.no_grad()
    def step(self, closure=None):
        self._owner._adam_step(self.param_groups[0])

    def load_state_dict(self, state_dict):
        """Restores step / exp_avg / exp_avg_sq INTO the packed buffers (the views in ``state`` stay linked)."""
        own = self._owner
        own._moments(own._require_cuda())
        params = [p for g in self.param_groups for p in g["params"]]
        for idx, st in state_dict.get("state", {}).items():
            p = params[int(idx)]
            if p in self.state and "exp_avg" in self.state[p]:
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                object.__setattr__(own, "_adam_t", int(st["step"]))
        for g_new, g in zip(state_dict.get("param_groups", []), self.param_groups):
            for k, v in g_new.items():
                if k != "params":
                    g[k] = v


class _FusedBase(nn.Module):
    """Shared engine plumbing of the two drop-in modules.

    Per-step training API: cached one-member EnsembleTrainers keyed by (combine, minibatch rows) -- a run with a
    ragged last batch alternates between two engines, nothing is rebuilt per step -- with the Adam moments kept in
    module-owned packed buffers.  Inference API (pred_recon / pred_latent / encode / decode): one cached engine per
    combine rule; rows are passed per call."""

    _loss_kind = "gauss_ll"
    _head_kind = None            # "regression" for cVAE_multimodal_regression, "endtoend" for e2e.cVAE_multimodal_endtoend
    _engine_flags = 0            # OR-ed into the per-step launch: _lib.TRAIN_FP32 selects the FP32 FFMA engine (bit-stable
                                 # trajectories; the default BF16x3 tensor-core engines meet the 1e-4 per-step bar)
    _opt_name = "optimizer1"     # attribute holding the fused Adam (the end-to-end class calls it `optimizer`)
    _ENGINE_KEYS = ("_engines", "_pending", "_last", "_views_cache", "_step_tensor", "_pending_fwd")

    def _head_kwargs(self):
        """make_arch / MemberSpec keywords of the supervised head (none for the normative models)."""
        return {}

    def _trainable(self):
        """(name in packed layout, Parameter) in optimizer1 order."""
        raise NotImplementedError

    def _trainable_named(self):
        return self._trainable()

    def _require_cuda(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("this cVAE runs only on CUDA (libnmb has no CPU fallback): call model.to('cuda')")
        return p.device

    # ---- engines -------------------------------------------------------------------------------------------
    def _cache(self):
        c = self.__dict__.get("_engines")
        if c is None:
            c = {}
            object.__setattr__(self, "_engines", c)
        return c

    def _make_engine(self, dev, dims, combine, rows, keep_grads, names=None, with_head=False):
        bufs = [torch.zeros((rows, _lib.packed_row_stride(int(d), self.c_dim)), dtype=torch.float32, device=dev)
                for d in dims]
        head = dict(self._head_kwargs()) if with_head else {}
        y = torch.zeros(rows, dtype=torch.float32, device=dev) if head else None
        spec = MemberSpec(dims, self._hidden, self.latent_dim, self.c_dim, bufs, combine=combine,
                          loss_kind=self._loss_kind, non_linear=self._non_linear, batch=rows, lr=self.learning_rate,
                          y=y, **head)
        eng = EnsembleTrainer([spec], device=dev, keep_grads=keep_grads)
        eng._rows_buf = bufs
        eng._y_buf = y
        return eng

    def _load_weights(self, eng, named=None, rename=None):
        """Current module parameters -> the engine's packed buffer (one multi-tensor copy)."""
        views = eng.__dict__.setdefault("_pviews", eng._views(0, eng.params))
        dst, src = [], []
        for name, p in (named or self._trainable_named()):
            k = rename(name) if rename else name
            if k is None or k not in views:           # e.g. the head's parameters in an engine built without the head
                continue
            dst.append(views[k]); src.append(p.detach().reshape(views[k].shape))
        for name, b in self._engine_buffers():              # e.g. BatchNorm running statistics (not parameters)
            if name in views:
                dst.append(views[name]); src.append(b.detach().to(torch.float32).reshape(views[name].shape))
        with torch.no_grad():
            torch._foreach_copy_(dst, src)

    def _engine_buffers(self):
        """(name in packed layout, buffer) pairs the engine reads AND updates (none for the normative models)."""
        return ()

    def _train_engine(self, xs, cs, combine):
        dev = self._require_cuda()
        rows = int(xs[0].shape[0])
        if rows > 256 or rows < 1:
            raise ValueError("a minibatch has 1..256 rows (train script :116)")
        key = ("train", combine.lower(), rows, str(dev)) + tuple(sorted(self._head_kwargs().items()))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, self._dims, combine, rows, keep_grads=True,
                                                         with_head=bool(self._head_kind))
        for dst, x, c in zip(eng._rows_buf, xs, cs):
            pack_rows(x.to(dev), c.to(dev), out=dst)
        self._load_weights(eng)
        return eng

    def _infer_engine(self, combine):
        dev = self._require_cuda()
        key = ("infer", combine.lower(), str(dev))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, self._dims, combine, 1, keep_grads=False)
        self._load_weights(eng)
        return eng

    def close(self):
        """Free every cached engine (they are also freed with the module)."""
        for eng in self._cache().values():
            eng.close()
        self._cache().clear()

    # ---- one fused forward + loss + backward launch ------------------------------------------------------------
    def _launch_step(self):
        eng, eps = self._pending
        eng.grads.zero_()
        loss4 = _lib.TRAIN_LOSS4 if self._head_kind else 0
        losses = eng.train_steps(1, eps=eps[None, None], record_losses=True,
                                 flags=_lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS | loss4
                                 | int(self._engine_flags))
        mu, lv, xr = eng.peek(0)
        gviews = eng.__dict__.setdefault("_gviews", eng._views(0, eng.grads))
        grads = [gviews[name].view(p.shape) for name, p in self._trainable_named()]       # views, nothing is copied
        lo = losses[0, 0]
        extra = [lo[3].reshape(())] if loss4 else []           # the head loss rides last (not differentiable by itself)
        return [lo[0].reshape(1), lo[1].reshape(()), lo[2].reshape(1), mu, lv] + list(xr) + extra, grads

    def _fused_forward(self, xs, cs, combine):
        eng = self._train_engine(xs, cs, combine)
        rows = xs[0].shape[0]
        # the reference draws eps with randn_like(mu) from the global generator (cVAE.py:420, 1132)
        eps = torch.randn((rows, self.latent_dim), device=xs[0].device if xs[0].is_cuda else eng.device,
                          dtype=torch.float32)
        object.__setattr__(self, "_pending", (eng, eps.to(eng.device)))
        params = [p for _, p in self._trainable_named()]
        return _FusedStep.apply(self, 0, *params)

    def _remember(self, fwd_rtn, losses, key):
        object.__setattr__(self, "_last", (fwd_rtn[key], losses))

    def _losses_of(self, fwd_rtn, key):
        last = self.__dict__.get("_last")
        if last is None:
            raise RuntimeError("loss_function called before forward")
        if not isinstance(fwd_rtn, dict) or fwd_rtn.get(key) is not last[0]:
            raise ValueError("fwd_rtn is not the result of this module's LAST forward pass: the fused kernel computes "
                             "the losses together with the forward pass, so only that pass can be scored")
        return dict(last[1])

    # ---- optimizer1.step() --------------------------------------------------------------------------------------
    def _moments(self, dev):
        """Module-owned packed exp_avg / exp_avg_sq (layout of nmb_arch_slots) and their per-parameter views."""
        m = self.__dict__.get("_adam_m")
        if m is None:
            arch = _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, "poe", self._loss_kind,
                                  self._non_linear, **self._head_kwargs())
            n = _lib.arch_param_count(arch)
            object.__setattr__(self, "_adam_m", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_v", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_t", 0)
        elif m.device != dev:                                  # the module was moved / unpickled on another device
            object.__setattr__(self, "_adam_m", self._adam_m.to(dev))
            object.__setattr__(self, "_adam_v", self._adam_v.to(dev))
            self.__dict__.pop("_views_cache", None)
        return self._adam_m, self._adam_v

    def _adam_step(self, group):
        last = self.__dict__.get("_pending")
        if last is None:
            raise RuntimeError("optimizer1.step() called before any forward pass")
        eng = last[0]
        named = list(self._trainable_named())
        m, v = self._moments(eng.device)
        pviews = eng.__dict__.setdefault("_pviews", eng._views(0, eng.params))
        gviews = eng.__dict__.setdefault("_gviews", eng._views(0, eng.grads))
        # gradients as the user left them (zero_grad / backward).  backward() hands out views of eng.grads, so in the
        # reference loop nothing is copied here; parameters without a gradient do not move (torch skips grad=None).
        for name, p in named:
            gv = gviews[name]
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad.reshape(gv.shape))
        self._load_weights(eng, named)
        t = self._adam_t + 1
        object.__setattr__(self, "_adam_t", t)
        b1, b2 = group["betas"]
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.nmb_adam_step(eng.params.data_ptr(), eng.grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                             eng.total_params, t, float(group["lr"]), float(b1), float(b2),
                                             float(group["eps"]), _stream_ptr(eng.device)), "nmb_adam_step")
        torch._foreach_copy_([p.data for _, p in named], [pviews[name].reshape(p.shape) for name, p in named])
        # torch.optim.Adam-shaped state (views of the packed moments), so state_dict() is meaningful
        vc = self.__dict__.get("_views_cache")
        if vc is None:
            vc = (eng._views(0, m), eng._views(0, v))
            object.__setattr__(self, "_views_cache", vc)
        st = getattr(self, self._opt_name).state
        step_t = self.__dict__.get("_step_tensor")
        if step_t is None or len(st) != len(named):          # built once; afterwards only the shared step counter moves
            step_t = torch.tensor(float(t))
            object.__setattr__(self, "_step_tensor", step_t)
            for name, p in named:
                st[p] = {"step": step_t, "exp_avg": vc[0][name].view(p.shape), "exp_avg_sq": vc[1][name].view(p.shape)}
        else:
            step_t.fill_(float(t))

    def __getstate__(self):      # torch.save(model): engines are not picklable and are rebuilt lazily
        d = self.__dict__.copy()
        for k in self._ENGINE_KEYS:
            d.pop(k, None)
        return d

    def __setstate__(self, state):
        """Also accepts a module pickled by the REFERENCE (``cVAE.cVAE`` / ``cVAE.cVAE_multimodal`` through the root
        ``cVAE.py`` shim): the private fields of the drop-in are derived from the reference's public attributes."""
        self.__dict__.update(state)
        if "_dims" not in self.__dict__:
            dims = self.__dict__.get("input_dim_list")
            self._dims = [int(d) for d in dims[: self.modalities]] if dims is not None else [int(self.input_dim)]
            self._hidden = [int(h) for h in list(self.hidden_dim)[:-1]]
            mods = self._modules
            enc = mods["encoder_list"][0] if "encoder_list" in mods else mods["encoder"]
            self._non_linear = bool(getattr(enc, "non_linear", True))
            opt = self.__dict__.get("optimizer1")
            if opt is not None and not isinstance(opt, _FusedAdam):       # the reference's torch.optim.Adam
                lr = opt.param_groups[0]["lr"]
                self.__dict__["optimizer1"] = _FusedAdam([p for _, p in self._trainable_named()], lr=lr, owner=self)


def _as_float_cuda(a, dev):
    if not torch.is_tensor(a):
        a = torch.as_tensor(np.asarray(getattr(a, "values", a)))      # DataFrame / ndarray / list
    return a.detach().to(device=dev, dtype=torch.float32)


def fuse_latent(mus, variances, combine, alphas=None):
    """``combine_latent`` (cVAE.py:1144-1164) with ProductOfExperts (:993-998), MixtureOfExperts (:1007-1011) and
    MoPoE (:1072-1083) as plain tensor ops on the caller's device (the training and scoring kernels carry their own
    fused implementation, csrc/nmb_fusion.cuh; this is the stand-alone method of the reference's API).
    mus, variances: [M, B, Z]."""
    kind = combine.lower()
    m = mus.shape[0]
    if kind == "poe":
        t = 1.0 / variances
        return torch.sum(mus * t, dim=0) / torch.sum(t, dim=0), 1.0 / torch.sum(t, dim=0)
    if kind == "gpoe":
        a = torch.softmax(torch.stack([p for p in alphas]), dim=0).reshape(m, 1, 1)
        return (torch.sum(mus * a / variances, dim=0) / torch.sum(a / variances, dim=0),
                1 / torch.sum(a / variances, dim=0))
    if kind == "moe":
        return torch.sum(mus, dim=0) / m, torch.sum(variances, dim=0) / m
    if kind == "mopoe":
        t = 1.0 / variances
        p_mu, p_var = torch.sum(mus * t, dim=0) / torch.sum(t, dim=0), 1.0 / torch.sum(t, dim=0)
        return (torch.sum(mus, dim=0) + p_mu) / (m + 1), (torch.sum(variances, dim=0) + p_var) / (m + 1)
    raise ValueError("No such combination method")


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
        # cVAE.py:412-413: the (never stepped) optimisers of the adversarial pieces, kept for attribute parity
        self.optimizer2 = optim.Adam(list(self.discriminator.parameters()), lr=learning_rate)
        self.optimizer3 = optim.Adam(list(self.encoder.parameters()), lr=learning_rate)

    def _trainable(self):
        for k, p in self.encoder.named_parameters():
            yield "encoder_list.0." + k, p
        for k, p in self.decoder.named_parameters():
            yield "decoder_list.0." + k, p

    # ---- the training-loop surface (cVAE.py:435-443, 491-504) ------------------------------------------
    def forward(self, x, c):
        self.zero_grad()
        total, kl, ll, mu, logvar, xr = self._fused_forward([x], [c], "poe")
        scale = self.decoder.logvar_out.detach().exp().pow(0.5)
        fwd = {"x_recon": Normal(loc=xr, scale=scale), "mu": mu, "logvar": logvar}
        self._remember(fwd, {"total": total, "kl": kl, "ll": ll}, "mu")
        return fwd

    def loss_function(self, x, fwd_rtn):
        """{'total','kl','ll'} of the forward pass that produced fwd_rtn (cVAE.py:491-504); any other fwd_rtn is
        refused with ValueError."""
        return self._losses_of(fwd_rtn, "mu")

    # ---- pieces (cVAE.py:415-433): inference-time calls into the kernels, no autograd graph -----------------
    def _packed(self, x, c, dev):
        c_t = _as_float_cuda(c, dev)
        x_t = _as_float_cuda(x, dev) if x is not None else torch.zeros((c_t.shape[0], self.input_dim), device=dev)
        return [pack_rows(x_t, c_t)]

    def encode(self, x, c):
        """(mu, logvar) of Encoder.forward (cVAE.py:415-416)."""
        dev = self._require_cuda()
        _, mu, lv = self._infer_engine("poe").reconstruct([self._packed(x, c, dev)], mode="mean", want_latent=True,
                                                           want_xhat=False)
        return mu[0], lv[0]

    def reparameterise(self, mu, logvar):
        """cVAE.py:418-421; eps from the global torch generator (randn_like), like the reference."""
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(mu)
        return mu + eps * std

    def decode(self, z, c):
        """Decoder.forward (cVAE.py:426-427): Normal(decoder_mean_layer(...), exp(logvar_out)^0.5)."""
        dev = self._require_cuda()
        xhat, _, _ = self._infer_engine("poe").reconstruct([self._packed(None, c, dev)], mode="decode",
                                                            eps=[_as_float_cuda(z, dev)])
        return Normal(loc=xhat[0][0], scale=self.decoder.logvar_out.detach().exp().pow(0.5))

    def calc_kl(self, mu, logvar):
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean(0)

    def calc_ll(self, x, x_recon):
        return compute_ll(x, x_recon)

    def _recon(self, x, c, mode, want_xhat=True):
        dev = self._require_cuda()
        out = self._infer_engine("poe").reconstruct([self._packed(x, c, dev)], mode=mode, want_latent=True,
                                                     want_xhat=want_xhat)
        torch.cuda.synchronize(dev)
        return out

    def pred_latent(self, x, c, DEVICE=None):
        """(mu, exp(logvar)) as numpy (cVAE.py:540-547)."""
        _, mu, lv = self._recon(x, c, "mean", want_xhat=False)
        return mu[0].cpu().numpy(), lv[0].exp().cpu().numpy()

    def pred_recon(self, x, c, DEVICE=None):
        """Decode the latent MEAN (cVAE.py:549-555)."""
        xhat, _, _ = self._recon(x, c, "mean")
        return xhat[0][0].cpu().numpy()


class cVAE_multimodal(_FusedBase):
    """M encoders + M decoders with latent fusion (cVAE.py:1087-1211)."""

    _rng_order = "cvae"          # alphas, encoders, decoders (cVAE.py:1107-1109)

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=3,
                 non_linear=False):
        super().__init__()
        self.input_dim_list = input_dim_list
        self.hidden_dim = hidden_dim + [latent_dim]
        self.latent_dim, self.c_dim, self.modalities, self.learning_rate = latent_dim, c_dim, modalities, learning_rate
        self._dims = [int(d) for d in input_dim_list[:modalities]]
        self._hidden, self._non_linear = list(hidden_dim), bool(non_linear)
        if self._rng_order == "cvae":
            self.alpha_m_list = nn.ParameterList(
                [nn.Parameter(torch.randn(1, requires_grad=True)) for _ in range(modalities)])
        self.encoder_list = nn.ModuleList(
            [Encoder(input_dim_list[i], self.hidden_dim, c_dim, non_linear) for i in range(modalities)])
        self.decoder_list = nn.ModuleList(
            [Decoder(input_dim_list[i], self.hidden_dim, c_dim, non_linear) for i in range(modalities)])
        if self._rng_order != "cvae":
            self.alpha_m_list = nn.ParameterList(
                [nn.Parameter(torch.randn(1, requires_grad=True)) for _ in range(modalities)])
        self._extra_init()
        self.optimizer1 = _FusedAdam(
            [p for m in self.encoder_list for p in m.parameters()]
            + [p for m in self.decoder_list for p in m.parameters()]
            + list(self.alpha_m_list.parameters()), lr=learning_rate, owner=self)

    def _extra_init(self):
        pass

    def _trainable(self):
        for i, m in enumerate(self.encoder_list):
            for k, p in m.named_parameters():
                yield f"encoder_list.{i}.{k}", p
        for i, m in enumerate(self.decoder_list):
            for k, p in m.named_parameters():
                yield f"decoder_list.{i}.{k}", p
        for i, p in enumerate(self.alpha_m_list):
            yield f"alpha_m_list.{i}", p

    # ---- the training-loop surface (cVAE.py:1166-1196) --------------------------------------------------
    def forward_multimodal(self, xes, cs, combine):
        self.zero_grad()
        _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, combine)   # ValueError if unknown
        outs = self._fused_forward(list(xes), list(cs), combine)
        total, kl, ll, mu, logvar = outs[:5]
        x_recons = [Normal(loc=outs[5 + i], scale=self.decoder_list[i].logvar_out.detach().exp().pow(0.5))
                    for i in range(self.modalities)]
        fwd = {"x_recons": x_recons, "mu_multimodal": mu, "logvar_multimodal": logvar}
        self._remember(fwd, {"total": total, "kl": kl, "ll": ll}, "mu_multimodal")
        return fwd

    def loss_function_multimodal(self, xes, fwd_rtn, labels=None):
        """sum_m (kl - ll_m) of the forward pass that produced fwd_rtn (cVAE.py:1187-1196); any other fwd_rtn is
        refused with ValueError.  `labels` is accepted (and ignored) for the nmmlp class, whose signature has it."""
        return self._losses_of(fwd_rtn, "mu_multimodal")

    # ---- pieces (cVAE.py:1127-1164): inference-time calls into the kernels, no autograd graph ------------------
    def _modality_engine(self, m):
        """One-modality engine holding encoder m / decoder m (for encode / decode of a single modality)."""
        dev = self._require_cuda()
        key = ("mod", m, str(dev))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, [self._dims[m]], "poe", 1, keep_grads=False)
        pre_e, pre_d = f"encoder_list.{m}.", f"decoder_list.{m}."

        def rename(name):
            if name.startswith(pre_e):
                return "encoder_list.0." + name[len(pre_e):]
            if name.startswith(pre_d):
                return "decoder_list.0." + name[len(pre_d):]
            return None
        self._load_weights(eng, rename=rename)
        return eng, dev

    def encode(self, x, c, m):
        """(mu, logvar) of encoder m (cVAE.py:1127-1128)."""
        eng, dev = self._modality_engine(m)
        xc = [pack_rows(_as_float_cuda(x, dev), _as_float_cuda(c, dev))]
        _, mu, lv = eng.reconstruct([xc], mode="mean", want_latent=True, want_xhat=False)
        return mu[0], lv[0]

    def reparameterise(self, mu, logvar):
        """cVAE.py:1130-1133."""
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(mu)
        return mu + eps * std

    def decode(self, z, c, m):
        """Normal of decoder m (cVAE.py:1135-1136)."""
        eng, dev = self._modality_engine(m)
        c_t = _as_float_cuda(c, dev)
        xc = [pack_rows(torch.zeros((c_t.shape[0], self._dims[m]), device=dev), c_t)]
        xhat, _, _ = eng.reconstruct([xc], mode="decode", eps=[_as_float_cuda(z, dev)])
        return Normal(loc=xhat[0][0], scale=self.decoder_list[m].logvar_out.detach().exp().pow(0.5))

    def product_of_experts(self, mus, variances):
        return fuse_latent(mus, variances, "poe")

    def mixture_of_experts(self, mus, variances):
        return fuse_latent(mus, variances, "moe")

    def mixture_of_product_of_experts(self, mus, variances):
        return fuse_latent(mus, variances, "mopoe")

    def combine_latent(self, mus, variances, combine):
        """cVAE.py:1144-1164 (case-insensitive; ValueError('No such combination method') otherwise)."""
        if self.modalities == 1 and self._rng_order != "nmmlp":
            return mus[0], variances[0]                      # :1145-1146
        return fuse_latent(mus, variances, combine, list(self.alpha_m_list))

    def calc_kl(self, mu, logvar):
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean(0)

    def calc_ll(self, x, x_recon):
        return compute_ll(x, x_recon)

    def pred_recon(self, xes, c, DEVICE=None, combine="poe"):
        """Test-time reconstruction; z is SAMPLED with torch.randn like the reference (cVAE.py:1198-1208)."""
        dev = self._require_cuda()
        c_t = _as_float_cuda(c, dev)
        xc = [pack_rows(_as_float_cuda(x, dev), c_t) for x in xes[: self.modalities]]
        eps = torch.randn((xc[0].shape[0], self.latent_dim), dtype=torch.float32)   # CPU generator, like :1207
        xhat, _, _ = self._infer_engine(combine).reconstruct([xc], mode="sample", eps=[eps.to(dev)])
        torch.cuda.synchronize(dev)
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
        model.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=False)
        return model.to(trainer.device)


class cVAE_multimodal_endtoend(cVAE_multimodal):
    """The model class defined inside multimodal_kfold_cvae_nmmlp.py (:39-245): same encoders / decoders / fusion,
    reconstruction term -MSE(mean) (:124-127), torch-RNG order encoders -> decoders -> alphas -> MLP (:57-83).  The MLP
    "diagnosis" head is constructed (it consumes RNG and appears in the state_dict) but, as in the reference, never
    called and not in optimizer1 (:92-98)."""

    _rng_order = "nmmlp"
    _loss_kind = "neg_mse"

    def _extra_init(self):
        self.total_input_size = sum(self._dims)
        self.mlp = nn.Sequential(nn.Linear(self.total_input_size, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(),
                                 nn.Linear(64, 1), nn.Sigmoid())
        self.criterion = nn.BCELoss()

    def calc_ll(self, x, x_recon):
        return -nn.MSELoss(reduction="mean")(x_recon.loc, x)

    def pred_recon(self, xes, c, DEVICE=None, combine="poe"):
        """nmmlp :212-233: like cVAE_multimodal.pred_recon but eps is drawn on the model's device (randn_like)."""
        dev = self._require_cuda()
        c_t = _as_float_cuda(c, dev)
        xc = [pack_rows(_as_float_cuda(x, dev), c_t) for x in xes[: self.modalities]]
        eps = torch.randn((xc[0].shape[0], self.latent_dim), dtype=torch.float32, device=dev)
        xhat, _, _ = self._infer_engine(combine).reconstruct([xc], mode="sample", eps=[eps])
        torch.cuda.synchronize(dev)
        return [t.cpu().numpy() for t in xhat[0]]


class cVAE_multimodal_regression(cVAE_multimodal):
    """``cVAE_multimodal_regression`` (cVAE.py:2211-2347): the multimodal cVAE plus a regressor MLP
    ``Linear(sum D, 128) ReLU Linear(128, 64) ReLU Linear(64, 1)`` on the concatenated residuals
    ``x_m - x_recon_m.loc`` (:2321-2325), trained jointly: ``total = sum_m (kl - ll_m) + lambda_reg * MSE(fi_pred,
    true_fi)`` (:2334-2347).  torch-RNG order of the constructor: encoders, decoders, alphas, regressor; Adam over
    encoders, decoders, regressor, alphas (:2260-2266).

    ``forward_multimodal`` runs the forward pass (encoders -> fusion -> z -> decoders -> residuals -> regressor) in one
    launch; ``loss_function_multimodal(xes, fwd_rtn, true_fi)`` -- the first moment the target is known -- runs the fused
    forward + losses + backward launch with the SAME eps draw, so ``losses['total'].backward()`` only hands out the
    gradients.  Runs on libnmb's generic tcgen05 engine (members with a head are not served by the pipelined kernel)."""

    _rng_order = "regression"
    _head_kind = "regression"
    _head_hidden = (128, 64)

    def _extra_init(self):
        widths = [sum(self._dims)] + list(self._head_hidden)
        layers = []
        for a, b in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        layers.append(nn.Linear(widths[-1], 1))
        self.regressor = nn.Sequential(*layers)
        self.mse_loss = nn.MSELoss()
        self._head_weight = 1.0

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=3, non_linear=False):
        super().__init__(input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate, modalities, non_linear)
        self.optimizer1 = _FusedAdam([p for _, p in self._trainable()], lr=learning_rate, owner=self)

    def _head_kwargs(self):
        return {"head": "regression", "head_hidden": tuple(self._head_hidden), "head_weight": float(self.__dict__.get("_head_weight", 1.0))}

    def _trainable(self):       # optimizer1 order (:2260-2266)
        for i, m in enumerate(self.encoder_list):
            for k, p in m.named_parameters():
                yield f"encoder_list.{i}.{k}", p
        for i, m in enumerate(self.decoder_list):
            for k, p in m.named_parameters():
                yield f"decoder_list.{i}.{k}", p
        for k, p in self.regressor.named_parameters():
            yield f"regressor.{k}", p
        for i, p in enumerate(self.alpha_m_list):
            yield f"alpha_m_list.{i}", p

    def forward_multimodal(self, xes, cs, combine):
        self.zero_grad()
        _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, combine)   # ValueError if unknown
        xs, cs = list(xes), list(cs)
        eng = self._train_engine(xs, cs, combine)
        rows = xs[0].shape[0]
        eps = torch.randn((rows, self.latent_dim), device=xs[0].device if xs[0].is_cuda else eng.device,
                          dtype=torch.float32).to(eng.device)             # reparameterise (:2284-2287)
        pred, xh, mu, lv = eng.head_predict([eng._rows_buf], mode="sample", eps=[eps], want_xhat=True, want_latent=True)
        x_recons = [Normal(loc=xh[0][i], scale=self.decoder_list[i].logvar_out.detach().exp().pow(0.5))
                    for i in range(self.modalities)]
        fwd = {"x_recons": x_recons, "mu_multimodal": mu[0], "logvar_multimodal": lv[0], "fi_pred": pred[0].view(-1, 1)}
        object.__setattr__(self, "_pending_fwd", (eng, eps, fwd, combine))
        return fwd

    def loss_function_multimodal(self, xes, fwd_rtn, true_fi, lambda_reg=1.0):
        """{'total', 'kl', 'll', 'regression'} of the forward pass that produced fwd_rtn (:2334-2347)."""
        last = self.__dict__.get("_pending_fwd")
        if last is None:
            raise RuntimeError("loss_function_multimodal called before forward_multimodal")
        eng, eps, fwd, combine = last
        if fwd_rtn is not fwd:
            raise ValueError("fwd_rtn is not the result of this module's LAST forward pass: the fused kernel computes "
                             "the losses together with the forward pass, so only that pass can be scored")
        if float(lambda_reg) != float(self.__dict__.get("_head_weight", 1.0)):
            # lambda_reg is part of the engine's architecture record: switch to (or build) the engine for this weight
            # and hand it the rows forward_multimodal packed
            self._head_weight = float(lambda_reg)
            rows = eng._rows_buf
            key = ("train", combine.lower(), int(rows[0].shape[0]), str(eng.device)) + tuple(sorted(self._head_kwargs().items()))
            eng2 = self._cache().get(key)
            if eng2 is None:
                eng2 = self._cache()[key] = self._make_engine(eng.device, self._dims, combine, int(rows[0].shape[0]),
                                                              keep_grads=True, with_head=True)
            for dst, src in zip(eng2._rows_buf, rows):
                dst.copy_(src)
            self._load_weights(eng2)
            eng = eng2
        eng._y_buf.copy_(torch.as_tensor(true_fi).to(device=eng.device, dtype=torch.float32).reshape(-1))
        object.__setattr__(self, "_pending", (eng, eps))
        outs = _FusedStep.apply(self, 0, *[p for _, p in self._trainable_named()])
        return {"total": outs[0], "kl": outs[1], "ll": outs[2], "regression": outs[-1]}


class mmJSD(cVAE_multimodal):
    """The ``mmJSD`` baseline of the model zoo (cVAE.py:1354-1452; ``-Model mmJSD`` in the train script :141-148;
    SURVEY 8 f4).  As written in the reference it is the multimodal cVAE with two fixed choices: ``forward_multimodal`` /
    ``pred_recon`` IGNORE ``combine`` and always fuse with the class's own product of experts (:1400-1403), and the
    Jensen-Shannon term is evaluated on M identical copies of the fused posterior (:1426) -- KL(p || p) -- so it is exactly 0
    with zero gradient (recorded: tests/golden/mmjsd_M3.npz ``jsd0``).  Same parameters, same torch-RNG order (alphas,
    encoders, decoders; the alphas never receive a gradient), same fused kernels."""

    def forward_multimodal(self, xes, cs, combine=None):
        return super().forward_multimodal(xes, cs, "poe")

    def pred_recon(self, xes, c, DEVICE=None, combine=None):
        return super().pred_recon(xes, c, DEVICE, "poe")

    def combine_latent(self, mus, logvars):
        """(mu, VARIANCE) of the product of the experts N(mus[m], exp(logvars[m])) (:1400-1403)."""
        return fuse_latent(mus, torch.exp(logvars), "poe")

    def reparameterize(self, mu, logvar):
        return self.reparameterise(mu, logvar)

    def multimodal_jsd(self, mus, logvars):
        """Mean pairwise KL(N_i || N_j) over the M (M - 1) / 2 pairs (:1405-1412)."""
        from torch.distributions import kl_divergence
        jsd, n = 0, len(mus)
        for i in range(n):
            for j in range(i + 1, n):
                jsd = jsd + kl_divergence(Normal(mus[i], torch.exp(0.5 * logvars[i])),
                                          Normal(mus[j], torch.exp(0.5 * logvars[j]))).mean()
        return jsd / (n * (n - 1) / 2)


class mvtCAE(cVAE_multimodal):
    """The ``mvtCAE`` baseline (cVAE.py:1754-1893; ``-Model mvtCAE``), AS WRITTEN in the reference: the
    cVAE_multimodal architecture and parameters, the fused variance clamped to >= 1e-6 (:1815), a 'poe' branch that hands
    variances to ProductOfExperts2 (which exponentiates them again, :1778 / :1800), and
    ``total = sum_m (kl + 1e-5 * ll_m + beta * tc)`` -- the log-likelihood with a plus sign (:1862), beta = 1e-4,
    tc = - sum_i mean_m logsumexp_b mu_m[b, i] (:1846-1853).  One fused launch per step (``NMB_FAMILY_MVTCAE``) on the
    generic engines; ``losses`` carries 'tc' as well."""

    _head_kind = "mvtcae"        # (selects the 4-wide loss record; there is no head)
    beta = 0.0001

    def _head_kwargs(self):
        return {"family": "mvtcae", "beta": float(self.beta)}

    def _make_engine(self, dev, dims, combine, rows, keep_grads, names=None, with_head=False):
        bufs = [torch.zeros((rows, _lib.packed_row_stride(int(d), self.c_dim)), dtype=torch.float32, device=dev) for d in dims]
        spec = MemberSpec(dims, self._hidden, self.latent_dim, self.c_dim, bufs, combine=combine, non_linear=self._non_linear,
                          batch=rows, lr=self.learning_rate, **self._head_kwargs())
        eng = EnsembleTrainer([spec], device=dev, keep_grads=keep_grads)
        eng._rows_buf = bufs
        return eng

    def _moments(self, dev):
        if self.__dict__.get("_adam_m") is None:
            n = _lib.arch_param_count(_lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, "poe", "gauss_ll",
                                                     self._non_linear, **self._head_kwargs()))
            object.__setattr__(self, "_adam_m", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_v", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_t", 0)
            return self._adam_m, self._adam_v
        return super()._moments(dev)

    def _infer_engine(self, combine):
        dev = self._require_cuda()
        key = ("infer", combine.lower(), str(dev))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, self._dims, combine, 1, keep_grads=False)
        self._load_weights(eng)
        return eng

    def forward_multimodal(self, xes, cs, combine):
        self.zero_grad()
        _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, combine)   # ValueError if unknown
        outs = self._fused_forward(list(xes), list(cs), combine)
        total, kl, ll, mu, logvar = outs[:5]
        m = self.modalities
        x_recons = [Normal(loc=outs[5 + i], scale=self.decoder_list[i].logvar_out.detach().exp().pow(0.5)) for i in range(m)]
        fwd = {"x_recons": x_recons, "mu_multimodal": mu, "logvar_multimodal": logvar, "qz_x": mu}
        self._remember(fwd, {"total": total, "kl": kl, "ll": ll, "tc": outs[5 + m]}, "mu_multimodal")
        return fwd

    def combine_latent(self, mus, variances, combine):
        """(mu, clamped variance) as the reference computes them (:1795-1816)."""
        if combine.lower() == "poe":
            t = 1.0 / torch.exp(variances)
            mu, var = torch.sum(mus * t, dim=0) / torch.sum(t, dim=0), torch.log(1.0 / torch.sum(t, dim=0))
        else:
            mu, var = fuse_latent(mus, variances, combine, list(self.alpha_m_list))
        return mu, torch.clamp(var, min=1e-6)
