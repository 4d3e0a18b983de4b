"""Drop-ins for the reference's end-to-end supervised model: ``Classifier`` and
``cVAE_multimodal_endtoend`` (v2, cVAE.py:2004-2207) and the program multimodal_kfold_cvae_nmpmcont.py.

Model: shared encoders -> PoE -> z -> a health and a disease decoder set + a classifier (Linear, BatchNorm1d, ReLU,
Dropout blocks) on z; loss = w_rec (rec_health + rec_disease) + w_kl kl + cross-entropy + w_con * contrastive hinge.
The whole step -- both decoder sets, batch-statistics BatchNorm, dropout, the five loss terms, backward, Adam and the
running-statistics update -- is ONE launch of libnmb's generic engines (``NMB_HEAD_ENDTOEND``); the program trains every
fold in one launch.  It runs on the FP32 FFMA engine (bit-stable trajectories: ReLU / BatchNorm / hinge decisions do not
depend on BF16x3 rounding; the tensor-core engine meets the same 1e-4 per-step bar, tests/test_gpu_e2e_head.py).

(This class is the one the reference's ``cVAE`` MODULE exports under this name.  The different class of the same name
that multimodal_kfold_cvae_nmmlp.py defines locally lives in ``.cVAE``.)
"""
from __future__ import annotations

import argparse
import random as rn
from os.path import join
from pathlib import Path

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import KFold
from sklearn.preprocessing import RobustScaler
from torch import nn
from torch.distributions import Normal

from . import _lib
from .cVAE import Decoder, Encoder, _FusedAdam, _FusedBase, _FusedStep
from .ensemble import EnsembleTrainer, MemberSpec, pack_rows
from .pipeline import covariate_onehots
from .utils import get_column_name, get_datasets_name, get_hc_label, load_dataset

LOSS_KEYS = ("total_loss", "recon_loss_health", "recon_loss_disease", "kl_loss", "classification_loss", "contrastive_loss")


class Classifier(nn.Module):
    """Parameter container with the reference's layout (cVAE.py:2004-2018): ``classifier.{4l}`` Linear, ``{4l+1}``
    BatchNorm1d, ReLU, Dropout, ..., final Linear.  It is evaluated inside the fused kernels (model.forward / predict)."""

    def __init__(self, latent_dim, classifier_layers, dropout_rate, num_classes=2):
        super().__init__()
        layers = []
        sizes = [latent_dim] + list(classifier_layers)
        for i in range(len(sizes) - 1):
            layers += [nn.Linear(sizes[i], sizes[i + 1]), nn.BatchNorm1d(sizes[i + 1]), nn.ReLU(), nn.Dropout(dropout_rate)]
        layers.append(nn.Linear(sizes[-1], num_classes))
        self.classifier = nn.Sequential(*layers)

    def forward(self, z):
        raise RuntimeError("the classifier runs inside libnmb's fused step: call cVAE_multimodal_endtoend.forward / predict")


class cVAE_multimodal_endtoend(_FusedBase):
    """cVAE.py:2021-2207.  ``forward`` (train mode) runs the fused forward pass with torch-drawn eps and dropout masks;
    ``loss_function(xes, fwd_rtn, labels, margin, weightcontrastive, weight_kl, weight_rec)`` -- the first moment the
    labels are known -- runs forward + losses + backward with the SAME draws and updates the BatchNorm running
    statistics once; ``losses['total_loss'].backward()`` hands out the gradients, ``optimizer.step()`` is the fused Adam.
    In eval mode ``forward`` samples z and uses the running statistics; ``predict`` classifies the fused mean."""

    _head_kind = "endtoend"
    _opt_name = "optimizer"

    def __init__(self, input_dim_list, hidden_dim, latent_dim, c_dim, learning_rate=0.0001, modalities=3, non_linear=False,
                 classifier_layers=[128, 64], dropout_rate=0.5, num_classes=2):
        super().__init__()
        if num_classes != 2:
            raise ValueError("the fused end-to-end step implements the reference's two-class setting (num_classes=2)")
        self.input_dim_list = input_dim_list
        self.hidden_dim = hidden_dim + [latent_dim]
        self.latent_dim, self.c_dim, self.modalities, self.learning_rate = latent_dim, c_dim, modalities, learning_rate
        self.non_linear, self.num_classes = non_linear, num_classes
        self._dims = [int(d) for d in input_dim_list[:modalities]]
        self._hidden, self._non_linear = list(hidden_dim), bool(non_linear)
        self._cls_layers, self._dropout = [int(w) for w in classifier_layers], float(dropout_rate)
        self._hp = dict(margin=1.0, w_contrastive=0.1, w_kl=0.1, w_rec=0.1)            # loss_function defaults (:2131)
        mk = lambda cls: nn.ModuleList([cls(input_dim_list[i], self.hidden_dim, c_dim, non_linear) for i in range(modalities)])
        self.encoder_list = mk(Encoder)                      # RNG order of the reference (:2042-2052)
        self.decoder_list_health = mk(Decoder)
        self.decoder_list_disease = mk(Decoder)
        self.classifier = Classifier(latent_dim, self._cls_layers, dropout_rate, num_classes)
        self.optimizer = _FusedAdam([p for _, p in self._trainable()], lr=learning_rate, owner=self)

    # ---- layout ------------------------------------------------------------------------------------------------
    def _trainable(self):          # optimizer order (:2055-2061)
        for pre in ("encoder_list", "decoder_list_health", "decoder_list_disease"):
            for i, m in enumerate(getattr(self, pre)):
                for k, p in m.named_parameters():
                    yield f"{pre}.{i}.{k}", p
        for k, p in self.classifier.named_parameters():
            yield f"classifier.{k}", p

    def _engine_buffers(self):
        return [(f"classifier.{k}", b) for k, b in self.classifier.named_buffers()]

    def _head_kwargs(self):
        hp = dict(self.__dict__.get("_hp") or {})
        hp["dropout"] = self._dropout
        return {"head": "endtoend", "head_hidden": tuple(self._cls_layers), "head_params": tuple(sorted(hp.items()))}

    def _make_engine(self, dev, dims, combine, rows, keep_grads, names=None, with_head=False):
        bufs = [torch.zeros((rows, _lib.packed_row_stride(int(d), self.c_dim)), dtype=torch.float32, device=dev) for d in dims]
        kw = {}
        if with_head:
            hk = self._head_kwargs()
            kw = dict(head="endtoend", head_hidden=hk["head_hidden"], head_params=dict(hk["head_params"]),
                      y=torch.zeros(rows, dtype=torch.float32, device=dev),
                      drop_keep=torch.ones((1, rows, sum(self._cls_layers)), dtype=torch.float32, device=dev))
        spec = MemberSpec(dims, self._hidden, self.latent_dim, self.c_dim, bufs, combine="poe", non_linear=self._non_linear,
                          batch=rows, lr=self.learning_rate, **kw)
        eng = EnsembleTrainer([spec], device=dev, keep_grads=keep_grads)
        eng._rows_buf, eng._y_buf, eng._keep_buf = bufs, kw.get("y"), kw.get("drop_keep")
        return eng

    def _moments(self, dev):
        m = self.__dict__.get("_adam_m")
        if m is None:
            hk = self._head_kwargs()
            arch = _lib.make_arch(self._dims, self._hidden, self.latent_dim, self.c_dim, "poe", "gauss_ll", self._non_linear,
                                  head="endtoend", head_hidden=hk["head_hidden"], head_params=dict(hk["head_params"]))
            n = _lib.arch_param_count(arch)
            object.__setattr__(self, "_adam_m", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_v", torch.zeros(n, dtype=torch.float32, device=dev))
            object.__setattr__(self, "_adam_t", 0)
            return self._adam_m, self._adam_v
        return super()._moments(dev)

    def _pull_buffers(self, eng):
        """BatchNorm running statistics the launch updated: engine -> module buffers."""
        views = eng.__dict__.setdefault("_pviews", eng._views(0, eng.params))
        with torch.no_grad():
            for name, b in self._engine_buffers():
                b.copy_(views[name].reshape(b.shape).to(b.dtype))

    # ---- one fused launch -----------------------------------------------------------------------------------------
    _FLAGS = _lib.TRAIN_FP32 | _lib.TRAIN_NO_ADAM | _lib.TRAIN_KEEP_ACTS | _lib.TRAIN_LOSS8

    def _launch(self, eng, eps, flags):
        losses = eng.train_steps(1, eps=eps[None, None], record_losses=True, flags=flags)
        mu, lv, xr = eng.peek(0)
        return losses[0, 0], mu, lv, xr, eng.peek_head(0)

    def _launch_step(self):        # called by _FusedStep.forward: the loss launch
        eng, eps = self._pending
        eng.grads.zero_()
        lo, mu, lv, xr, logits = self._launch(eng, eps, self._FLAGS | _lib.TRAIN_WRITE_GRADS)
        self._pull_buffers(eng)
        gviews = eng.__dict__.setdefault("_gviews", eng._views(0, eng.grads))
        grads = [gviews[name].view(p.shape) for name, p in self._trainable_named()]
        # outputs 0..2 follow the (total, kl, ll) convention of _FusedStep; the rest is not differentiable by itself
        return [lo[0].reshape(()), lo[1].reshape(()), lo[2].reshape(()), lo[3].reshape(()), lo[4].reshape(()),
                lo[5].reshape(()), lo[6].reshape(())], grads

    def forward(self, xes, cs):
        xs, cs = list(xes), list(cs)
        if not self.training:
            return self._forward_eval(xs, cs)
        self.zero_grad()
        eng = self._train_engine(xs, cs, "poe")
        rows = xs[0].shape[0]
        dev = eng.device
        eps = torch.randn((rows, self.latent_dim), device=dev, dtype=torch.float32)                  # reparameterise (:2076-2079)
        eng._keep_buf[0].copy_((torch.rand((rows, sum(self._cls_layers)), device=dev) >= self._dropout).float())   # nn.Dropout
        eng._y_buf.zero_()
        _, mu, lv, xr, logits = self._launch(eng, eps, self._FLAGS | _lib.TRAIN_NO_STATS)
        fwd = self._fwd_dict(mu, lv, xr, logits)
        object.__setattr__(self, "_pending_fwd", (eng, eps, fwd))
        return fwd

    def _fwd_dict(self, mu, lv, xr, logits):
        m = self.modalities
        scale = lambda dec: dec.logvar_out.detach().exp().pow(0.5)
        return {"x_recons_health": [Normal(loc=xr[i], scale=scale(self.decoder_list_health[i])) for i in range(m)],
                "x_recons_disease": [Normal(loc=xr[m + i], scale=scale(self.decoder_list_disease[i])) for i in range(m)],
                "mu": mu, "logvar": lv, "logits": logits}

    def _forward_eval(self, xs, cs):
        """model.eval(): z is still sampled (:2111), BatchNorm uses its running statistics, dropout is off."""
        eng = self._train_engine(xs, cs, "poe")
        eps = torch.randn((xs[0].shape[0], self.latent_dim), device=eng.device, dtype=torch.float32)
        logits, xh, mu, lv = eng.head_predict([eng._rows_buf], mode="sample", eps=[eps], engine="fp32", want_xhat=True,
                                              want_latent=True)
        return self._fwd_dict(mu[0], lv[0], xh[0], logits[0])

    def loss_function(self, xes, fwd_rtn, labels, margin=1.0, weightcontrastive=0.1, weight_kl=0.1, weight_rec=0.1):
        last = self.__dict__.get("_pending_fwd")
        if last is None:
            raise RuntimeError("loss_function called before forward (in training mode)")
        eng, eps, fwd = last
        if fwd_rtn is not fwd:
            raise ValueError("fwd_rtn is not the result of this module's LAST forward pass: the fused kernel computes "
                             "the losses together with the forward pass, so only that pass can be scored")
        hp = dict(margin=float(margin), w_contrastive=float(weightcontrastive), w_kl=float(weight_kl), w_rec=float(weight_rec))
        if hp != self._hp:           # the loss weights are part of the engine's architecture record: switch engines
            self._hp = hp
            rows = eng._rows_buf
            key = ("train", "poe", int(rows[0].shape[0]), str(eng.device)) + tuple(sorted(self._head_kwargs().items()))
            eng2 = self._cache().get(key)
            if eng2 is None:
                eng2 = self._cache()[key] = self._make_engine(eng.device, self._dims, "poe", int(rows[0].shape[0]),
                                                              keep_grads=True, with_head=True)
            for dst, src in zip(eng2._rows_buf, rows):
                dst.copy_(src)
            eng2._keep_buf.copy_(eng._keep_buf)
            self._load_weights(eng2)
            eng = eng2
        eng._y_buf.copy_(torch.as_tensor(labels).to(device=eng.device, dtype=torch.float32).reshape(-1))
        object.__setattr__(self, "_pending", (eng, eps))
        o = _FusedStep.apply(self, 0, *[p for _, p in self._trainable_named()])
        # LOSS8 order: total, kl, ll, ce, rec_health, rec_disease, contrastive
        return {"total_loss": o[0], "recon_loss_health": o[4], "recon_loss_disease": o[5], "kl_loss": o[1],
                "classification_loss": o[3], "contrastive_loss": o[6]}

    def predict(self, xes, cs):
        """logits = classifier(mu_combined) (:2198-2203); call under model.eval() like the reference's evaluate()."""
        xs, cs = list(xes), list(cs)
        dev = self._require_cuda()
        from .cVAE import _as_float_cuda
        xc = [pack_rows(_as_float_cuda(x, dev), _as_float_cuda(c, dev)) for x, c in zip(xs, cs)]
        key = ("predict", str(dev))
        eng = self._cache().get(key)
        if eng is None:
            eng = self._cache()[key] = self._make_engine(dev, self._dims, "poe", 1, keep_grads=False, with_head=True)
        self._load_weights(eng)
        return eng.head_predict([xc], mode="mean", engine="fp32")[0]

    def calc_kl(self, mu, logvar):
        return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1).mean()

    def compute_deviation(self, x, x_recon):
        return ((x - x_recon.mean) ** 2).mean(dim=1)


# ---- the program: multimodal_kfold_cvae_nmpmcont.py ---------------------------------------------------------------------
def generate_kfold_ids_endtoend(root, HC_group, other_group, oversample_percentage=1, n_splits=5, random_state=42):
    """utils.py:19-42: KFold over HC + other together, bootstrap of the training ids with the legacy numpy RNG; written to
    outputs/kfold_analysis_endtoend/."""
    kf = KFold(n_splits=n_splits, shuffle=True, random_state=random_state)
    kfold_dir = Path(root) / "outputs" / "kfold_analysis_endtoend"
    kfold_dir.mkdir(parents=True, exist_ok=True)
    all_group = pd.concat([HC_group, other_group])
    for fold, (train_idx, test_idx) in enumerate(kf.split(all_group)):
        train_ids = all_group.iloc[train_idx]["IID"]
        test_ids = all_group.iloc[test_idx]["IID"]
        over = np.random.choice(train_ids, size=int(len(train_ids) * oversample_percentage), replace=True)
        pd.DataFrame({"IID": over}).to_csv(kfold_dir / f"train_ids_{fold:03d}.csv", index=False)
        test_ids.to_csv(kfold_dir / f"test_ids_{fold:03d}.csv", index=False)
    return kfold_dir


def process_dataset(df, columns, scaler=None, fit_scaler=False, hc_label=None):
    """nmpmcont :75-123: RobustScaler, rank-quantile one-hot covariates (27 + 2), labels healthy = 0 / disease = 1."""
    data = df[columns].values
    scaler = scaler or RobustScaler()
    data = scaler.fit_transform(data) if fit_scaler else scaler.transform(data)
    labels = (df["DIA"].to_numpy() != hc_label).astype(np.int64)
    return np.asarray(data, dtype=np.float64), covariate_onehots(df), labels, scaler


def classification_metrics(labels, preds):
    """evaluate() :47-71 (sklearn's definitions on the hard predictions)."""
    from sklearn.metrics import accuracy_score, confusion_matrix, f1_score, recall_score, roc_auc_score
    try:
        auroc = roc_auc_score(labels, preds)
    except ValueError:
        auroc = float("nan")
    tn, fp, fn, tp = confusion_matrix(labels, preds, labels=[0, 1]).ravel()
    return {"accuracy": accuracy_score(labels, preds), "auroc": auroc, "sensitivity": recall_score(labels, preds, zero_division=0),
            "specificity": tn / (tn + fp) if tn + fp else float("nan"), "f1_score": f1_score(labels, preds, zero_division=0)}


def build_parser():
    p = argparse.ArgumentParser()                      # the reference's flags (:341-443)
    p.add_argument("-R", "--dataset_resourse", dest="dataset_resourse", type=str)
    p.add_argument("-H", "--hz_para_list", dest="hz_para_list", nargs="+", type=int)
    p.add_argument("-C", "--combine", dest="combine", type=str)
    p.add_argument("-P", "--procedure", dest="procedure", type=str)
    p.add_argument("-E", "--epochs", dest="epochs", type=int)
    p.add_argument("-K", "--n_splits", dest="n_splits", type=int, default=5)
    p.add_argument("-O", "--oversample_percentage", dest="oversample_percentage", type=float, default=1)
    p.add_argument("-Model", "--model", default="cVAE_multimodal", dest="model", type=str)
    p.add_argument("-SingleModality", "--single_modality", dest="single_modality", default=None, type=str)
    p.add_argument("-Baselearningrate", "--base_learning_rate", dest="base_learning_rate", type=float, default=0.0001)
    p.add_argument("-Maxlearningrate", "--max_learning_rate", dest="max_learning_rate", type=float, default=0.005)
    p.add_argument("-Learningrateclassifier", "--learning_rate_classifier", dest="learning_rate_classifier", type=float, default=0.001)
    p.add_argument("-Margin", "--margin", dest="margin", type=float, default=1)
    p.add_argument("-Weightcontrastive", "--weightcontrastive", dest="weightcontrastive", type=float, default=1)
    p.add_argument("-Weightkl", "--weight_kl", dest="weight_kl", type=float, default=1)
    p.add_argument("-Weightrec", "--weight_rec", dest="weight_rec", type=float, default=1)
    p.add_argument("-Dropout", "--dropout", dest="dropout", type=float, default=0.5)
    p.add_argument("-Layers", "--layers", dest="layers", nargs="+", default=[128, 64, 32], type=int)
    return p


def fill_defaults(args):               # :447-466
    if args.hz_para_list is None:
        args.hz_para_list = [110, 110, 10]
    if args.procedure is None:
        args.procedure = "SE-MoE"
    if args.combine is None:
        args.combine = args.procedure.split("-")[1]
    if args.dataset_resourse is None:
        args.dataset_resourse = "ADNI"
    if args.epochs is None:
        args.epochs = 200
    return args


def main(args, root=None, debug=None):
    """Every fold of the k-fold run trains in ONE fused launch; evaluation is one predict launch.

    Kept from the reference, quirks included: the fold ids are generated into outputs/kfold_analysis_endtoend but READ
    from outputs/kfold_analysis (:170-171 vs utils.py:21) -- here the freshly written ids are used when that directory
    has none; ``torch.manual_seed(42)`` right before each model, so every fold starts from the same weights (:196);
    ``optimizer.lr = clr`` never reaches Adam, which runs at the constructor's 1e-4 (:233); weight_kl / weight_rec /
    dropout of the command line are not passed on (0.1 / 0.1 / 0.5 apply, :241, :217)."""
    if not torch.cuda.is_available():
        raise RuntimeError("the end-to-end program needs a CUDA device (libnmb has no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    args = fill_defaults(args)
    root = Path(root or Path.cwd())
    model_dir = root / "outputs" / "kfold_analysis" / "supervised_cvae"
    model_dir.mkdir(parents=True, exist_ok=True)
    np.random.seed(42)
    rn.seed(42)
    names = get_datasets_name(args.dataset_resourse, args.procedure)
    participants_path = root / "data" / args.dataset_resourse / "y.csv"
    ids_df = pd.read_csv(participants_path)
    hc_label = get_hc_label(args.dataset_resourse)
    new_dir = generate_kfold_ids_endtoend(root, ids_df[ids_df["DIA"] == hc_label], ids_df[ids_df["DIA"] != hc_label],
                                          args.oversample_percentage, args.n_splits)
    ref_dir = root / "outputs" / "kfold_analysis"
    ids_dir = ref_dir if (ref_dir / "train_ids_000.csv").exists() else new_dir
    h_dim, z_dim = list(args.hz_para_list[:-1]), int(args.hz_para_list[-1])
    specs, tests, models = [], [], []
    for fold in range(args.n_splits):
        (model_dir / f"{fold:03d}").mkdir(exist_ok=True)
        xc_tr, xc_te, dims, lab_tr, lab_te, scaler = [], [], [], None, None, None
        for name in names:
            cols = get_column_name(args.dataset_resourse, name)
            feat = root / "data" / args.dataset_resourse / f"{name}.csv"
            tr = load_dataset(participants_path, ids_dir / f"train_ids_{fold:03d}.csv", feat)
            te = load_dataset(participants_path, ids_dir / f"test_ids_{fold:03d}.csv", feat)
            x_tr, c_tr, l_tr, scaler = process_dataset(tr, cols, scaler, True, hc_label)
            x_te, c_te, l_te, _ = process_dataset(te, cols, scaler, False, hc_label)
            if lab_tr is None:
                lab_tr, lab_te = l_tr, l_te                 # labels = label_list[0] (:236)
            xc_tr.append(pack_rows(torch.from_numpy(x_tr.astype(np.float32)).to(dev), torch.from_numpy(c_tr).to(dev)))
            xc_te.append(pack_rows(torch.from_numpy(x_te.astype(np.float32)).to(dev), torch.from_numpy(c_te).to(dev)))
            dims.append(x_tr.shape[1])
        torch.manual_seed(42)
        model = cVAE_multimodal_endtoend(input_dim_list=dims, hidden_dim=h_dim, latent_dim=z_dim, c_dim=c_tr.shape[1],
                                         modalities=len(names), non_linear=True, classifier_layers=list(args.layers),
                                         dropout_rate=0.5, num_classes=2)
        hp = dict(margin=float(args.margin), w_contrastive=float(args.weightcontrastive), w_kl=0.1, w_rec=0.1, dropout=0.5)
        specs.append(MemberSpec(dims, h_dim, z_dim, c_tr.shape[1], xc_tr, combine="poe", batch=256, seed=4242 + fold,
                                lr=model.learning_rate, state_dict=model.state_dict(), head="endtoend",
                                head_hidden=list(args.layers), head_params=hp,
                                y=torch.from_numpy(lab_tr.astype(np.float32)).to(dev), tag=fold))
        tests.append((xc_te, lab_te))
        models.append(model)
        if debug is not None:
            debug.setdefault("folds", []).append({"init": {k: v.clone() for k, v in model.state_dict().items()}, "dims": dims,
                                                  "xc_train": [t.cpu().numpy() for t in xc_tr], "labels": lab_tr,
                                                  "xc_test": [t.cpu().numpy() for t in xc_te], "seed": 4242 + fold, "hp": hp})
    trainer = EnsembleTrainer(specs, device=dev)
    losses = trainer.train_epochs(args.epochs, record_losses=True, flags=_lib.TRAIN_FP32 | _lib.TRAIN_LOSS8)
    logits = trainer.head_predict([t[0] for t in tests], mode="mean", engine="fp32")
    torch.cuda.synchronize(dev)
    if debug is not None:
        debug["losses"], debug["logits"] = losses.cpu().numpy(), [l.cpu().numpy() for l in logits]
        debug["steps_per_epoch"] = list(trainer.steps_per_epoch)
    all_metrics = []
    for fold in range(args.n_splits):
        spe = trainer.steps_per_epoch[fold]
        lo = losses[fold].cpu().numpy()
        for epoch in (0, args.epochs - 1):               # the reference prints batch 0 of every epoch (:249-251)
            r = lo[epoch * spe]
            print(f"Train Epoch:{epoch} Train batch: 0 total_loss: {r[0]:.3f}, recon_loss_health: {r[4]:.3f}, "
                  f"recon_loss_disease: {r[5]:.3f}, kl_loss: {r[1]:.3f}, classification_loss: {r[3]:.3f}, "
                  f"contrastive_loss: {r[6]:.3f}")
        model = models[fold]
        model.load_state_dict({k: v.cpu() for k, v in trainer.state_dict(fold).items()}, strict=True)
        torch.save(model, join(model_dir / f"{fold:03d}", "cVAE_model.pkl"))
        preds = logits[fold].argmax(dim=1).cpu().numpy()
        metrics = classification_metrics(tests[fold][1], preds)
        print(f"Fold {fold} metrics:")
        print(metrics)
        all_metrics.append(metrics)
    trainer.close()
    df = pd.DataFrame(all_metrics)
    print(df.mean())
    print(df.std())
    with open(join(root, "results_endtoend.csv"), "a") as f:            # :325-338
        f.write(str(args) + "\n")
        for metric in df.mean().index:
            f.write(f"{metric} ${df.mean()[metric]:.3f} \\pm {df.std()[metric]:.3f}$\n")
        f.write("\n\n\n")
    return all_metrics


def cli_main(argv=None):
    main(build_parser().parse_args(argv))
