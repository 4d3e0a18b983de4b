"""Drop-in for multimodal_kfold_train_cvae_supervised_regression.py: k-fold training + evaluation of
``cVAE_multimodal_regression`` -- the multimodal cVAE with an FI regressor on the reconstruction residuals.

The reference trains the folds one after the other, each with ``epochs x ceil(n / batch)`` Python-level steps
(:55-131).  Here every fold is one member of an ensemble with a regression head (``MemberSpec(head="regression")``) and
ALL folds train in ONE launch of libnmb's generic tcgen05 engine; the evaluation pass (:133-152) is one
``nmb_ensemble_head_predict`` launch.  What the host keeps doing is what is integer / RNG work in the reference:

  * KFold(shuffle=True, random_state=42) over y.csv (:53-55), the pandas merges and RobustScaler (:66-82);
  * the model constructors under ``torch.manual_seed(42)`` (:40, :102-110) -- seed-exact initial weights;
  * the loaders' shuffling: ``DataLoader(shuffle=True)`` per modality (:94) draws an independent permutation per
    modality and epoch from the CPU generator (the modalities of one minibatch are therefore NOT the same subjects;
    the target follows modality 0, :125).  ``loader_orders`` replays exactly those draws with real DataLoaders over row
    indices and hands the permutations to the kernel (``NmbMember.row_order``).

eps comes from the in-kernel Philox stream instead of torch's CUDA generator (as everywhere in this package).
Outputs as in the reference: regression_outputs/fold_{f}_{pred,true}.npy, deviation_fold_{f}_{modality}_roiwise.csv
(:170-199); the scatter plot (:159-168) is written only if matplotlib is importable.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import KFold
from sklearn.preprocessing import RobustScaler

from .cVAE import cVAE_multimodal_regression
from .ensemble import EnsembleTrainer, MemberSpec, pack_rows
from .utils import get_column_name, get_datasets_name


def evaluate_regression(y_true, y_pred):
    """RMSE / MAE / R2 / MAPE of the reference (:30-35; sklearn's definitions written out)."""
    y_true, y_pred = np.asarray(y_true, dtype=np.float64), np.asarray(y_pred, dtype=np.float64)
    err = y_true - y_pred
    ss_res, ss_tot = float(np.sum(err ** 2)), float(np.sum((y_true - y_true.mean()) ** 2))
    return {"RMSE": float(np.sqrt(np.mean(err ** 2))), "MAE": float(np.mean(np.abs(err))),
            "R2": 1.0 - ss_res / ss_tot if ss_tot > 0 else float("nan"),
            "MAPE": float(np.mean(np.abs(err / (y_true + 1e-6))) * 100)}


def loader_orders(n: int, batch_size: int, epochs: int, n_mod: int) -> np.ndarray:
    """[epochs, n_mod, n] int32: row yielded at each position by ``zip(*[DataLoader(ds_m, batch_size, shuffle=True)])``
    epoch after epoch (:94, :121-122), consuming torch's CPU generator exactly as that loop does."""
    index = torch.arange(n)
    loaders = [torch.utils.data.DataLoader(index, batch_size=batch_size, shuffle=True) for _ in range(n_mod)]
    out = np.empty((epochs, n_mod, n), dtype=np.int32)
    for ep in range(epochs):
        parts = [[] for _ in range(n_mod)]
        for batches in zip(*loaders):
            for m, b in enumerate(batches):
                parts[m].append(b.numpy())
        for m in range(n_mod):
            out[ep, m] = np.concatenate(parts[m])
    return out


def _consume_eval_loaders(n: int, batch_size: int, n_mod: int):
    """The evaluation loop's ``zip(*generator_test_list)`` (:140) creates one loader iterator per modality; each draws a
    base seed from the CPU generator even with shuffle=False.  Replayed so that the next fold's constructor sees the
    generator state the reference would give it."""
    loaders = [torch.utils.data.DataLoader(torch.arange(n), batch_size=batch_size, shuffle=False) for _ in range(n_mod)]
    for _ in zip(*loaders):
        pass


def build_parser():
    p = argparse.ArgumentParser()                      # the reference's flags (:199-207)
    p.add_argument("-R", "--dataset_resourse", type=str, default="ADNI")
    p.add_argument("-H", "--hz_para_list", nargs="+", type=int, default=[110, 110, 10])
    p.add_argument("-C", "--combine", type=str, default="gpoe")
    p.add_argument("-P", "--procedure", type=str, default="UCA-gPoE")
    p.add_argument("-E", "--epochs", type=int, default=500)
    p.add_argument("-K", "--n_splits", type=int, default=5)
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("-BaseLR", "--base_learning_rate", type=float, default=0.0001)
    return p


def train_and_test(args, root=None, debug=None):
    """All folds in one fused training launch + one evaluation launch.  Returns the per-fold score dicts.
    debug: optional dict that receives what the parity test replays through the oracle (per fold: initial weights,
    scaled inputs, loader permutations, per-step losses, predictions)."""
    if not torch.cuda.is_available():
        raise RuntimeError("the regression program needs a CUDA device (libnmb has no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    root = Path(root or Path.cwd())
    torch.manual_seed(42)                               # :39-41
    np.random.seed(42)
    output_dir = root / "regression_outputs"
    output_dir.mkdir(exist_ok=True)
    names = get_datasets_name(args.dataset_resourse, args.procedure)
    participants_path = root / "data" / args.dataset_resourse / "y.csv"
    ids_df = pd.read_csv(participants_path)
    demo_df = pd.read_csv(participants_path)
    tables = {n_: pd.read_csv(root / "data" / args.dataset_resourse / f"{n_}.csv") for n_ in names}
    cols = {n_: get_column_name(args.dataset_resourse, n_) for n_ in names}
    h_dim, z_dim = list(args.hz_para_list[:-1]), int(args.hz_para_list[-1])
    kf = KFold(n_splits=args.n_splits, shuffle=True, random_state=42)
    specs, tests, models = [], [], []
    for fold, (train_idx, test_idx) in enumerate(kf.split(ids_df)):
        train_ids = ids_df.iloc[train_idx]["IID"].tolist()
        test_ids = ids_df.iloc[test_idx]["IID"].tolist()
        xc_train, xc_test, dims, fi_train, fi_test = [], [], [], None, None
        for m, name in enumerate(names):
            mod = tables[name]
            tr = pd.merge(mod[mod["IID"].isin(train_ids)], demo_df, on="IID")             # :70-71
            te = pd.merge(mod[mod["IID"].isin(test_ids)], demo_df, on="IID")
            scaler = RobustScaler()
            x_tr = scaler.fit_transform(tr[cols[name]].values).astype(np.float32)         # :76-78, :91
            x_te = scaler.transform(te[cols[name]].values).astype(np.float32)
            c_tr = tr[["AGE", "PTGENDER"]].values.astype(np.float32)                      # raw covariates (:80-81)
            c_te = te[["AGE", "PTGENDER"]].values.astype(np.float32)
            if m == 0:                                                                    # the target of loader 0 (:125)
                fi_train = tr["FI"].values.astype(np.float32)
                fi_test = te["FI"].values.astype(np.float32).reshape(-1, 1)
            xc_train.append(pack_rows(torch.from_numpy(x_tr).to(dev), torch.from_numpy(c_tr).to(dev)))
            xc_test.append(pack_rows(torch.from_numpy(x_te).to(dev), torch.from_numpy(c_te).to(dev)))
            dims.append(x_tr.shape[1])
        n = int(xc_train[0].shape[0])
        model = cVAE_multimodal_regression(input_dim_list=dims, hidden_dim=h_dim, latent_dim=z_dim, c_dim=2,
                                           learning_rate=args.base_learning_rate, modalities=len(names), non_linear=True)
        order = loader_orders(n, args.batch_size, args.epochs, len(names))
        _consume_eval_loaders(int(xc_test[0].shape[0]), args.batch_size, len(names))
        specs.append(MemberSpec(dims, h_dim, z_dim, 2, xc_train, combine=args.combine, batch=args.batch_size,
                                seed=4242 + fold, lr=args.base_learning_rate, state_dict=model.state_dict(),
                                head="regression", y=torch.from_numpy(fi_train).to(dev),
                                row_order=torch.from_numpy(order).to(dev), tag=fold))
        tests.append((xc_test, fi_test))
        models.append(model)
        if debug is not None:
            debug.setdefault("folds", []).append({
                "init": {k: v.clone() for k, v in model.state_dict().items()}, "order": order, "fi": fi_train,
                "xc_train": [t.cpu().numpy() for t in xc_train], "xc_test": [t.cpu().numpy() for t in xc_test],
                "dims": list(dims), "seed": 4242 + fold})
    trainer = EnsembleTrainer(specs, device=dev)
    losses = trainer.train_epochs(args.epochs, record_losses=True, flags=_loss4())
    preds = trainer.head_predict([t[0] for t in tests], mode="sample")                    # z is sampled at test time too
    torch.cuda.synchronize(dev)
    if debug is not None:
        debug["losses"] = losses.cpu().numpy()
        debug["preds"] = [p.cpu().numpy() for p in preds]
        debug["steps_per_epoch"] = list(trainer.steps_per_epoch)
    all_scores = []
    for fold in range(len(specs)):
        spe = trainer.steps_per_epoch[fold]
        last = losses[fold, args.epochs * spe - 1].cpu().numpy()
        print(f"[Fold {fold}][Epoch {args.epochs - 1}] Loss: {last[0]:.4f}, FI MSE: {last[3]:.4f}")      # :132-133
        pred = preds[fold].cpu().numpy().reshape(-1, 1)
        true = tests[fold][1]
        np.save(output_dir / f"fold_{fold}_pred.npy", pred)
        np.save(output_dir / f"fold_{fold}_true.npy", true)
        scores = evaluate_regression(true, pred)
        print(f"[Fold {fold}] RMSE: {scores['RMSE']:.4f}, MAE: {scores['MAE']:.4f}, R²: {scores['R2']:.4f}, "
              f"MAPE: {scores['MAPE']:.2f}%")
        all_scores.append(scores)
        _scatter(true, pred, fold, output_dir)
        model = models[fold]
        model.load_state_dict({k: v.cpu() for k, v in trainer.state_dict(fold).items()}, strict=False)
        model.to(dev).eval()
        all_ids = ids_df["IID"].tolist()
        for m, name in enumerate(names):                                                  # ROI-wise deviations (:170-199)
            mod = tables[name]
            full = pd.merge(mod[mod["IID"].isin(all_ids)], demo_df, on="IID")
            x = RobustScaler().fit_transform(full[cols[name]].values.astype(np.float32))
            x_t = torch.tensor(x, dtype=torch.float32).to(dev)
            c_t = torch.tensor(full[["AGE", "PTGENDER"]].values.astype(np.float32)).to(dev)
            with torch.no_grad():
                mu, logvar = model.encode(x_t, c_t, m)
                z = model.reparameterise(mu, logvar)
                x_recon = model.decode(z, c_t, m).loc
                dev_roi = ((x_t - x_recon) ** 2).cpu().numpy()
            out = pd.DataFrame(dev_roi, columns=[f"ROI_{i}" for i in range(dev_roi.shape[1])])
            out.insert(0, "IID", full["IID"].tolist())
            from .cli import write_csv
            write_csv(out, output_dir / f"deviation_fold_{fold}_{name}_roiwise.csv")
        model.close()
    trainer.close()
    print("Training & evaluation complete.")
    return all_scores


def _loss4():
    from . import _lib
    return _lib.TRAIN_LOSS4


def _scatter(true, pred, fold, output_dir):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    plt.figure(figsize=(6, 6))
    plt.scatter(true, pred, alpha=0.5)
    plt.plot([true.min(), true.max()], [true.min(), true.max()], "r--")
    plt.xlabel("True FI"); plt.ylabel("Predicted FI"); plt.title(f"Fold {fold} - FI Prediction"); plt.grid(True)
    plt.savefig(output_dir / f"fold_{fold}_scatter.png")
    plt.close()


def main(argv=None):
    train_and_test(build_parser().parse_args(argv))
