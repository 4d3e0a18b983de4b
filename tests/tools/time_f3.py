"""Dev tool: the supervised regression head timed -- K folds of cVAE_multimodal_regression (3 x 116 ROIs,
hidden 110/110, latent 10, batch 128, shuffling loaders) trained in one launch of the generic engines, against the
oracle's restatement of the reference loop (torch CPU, all host threads) on the same fold."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
from multi_modal_normative_modeling_b200.regression import loader_orders

folds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n, b, dims = 800, 128, [116, 116, 116]
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
torch.manual_seed(0)
specs = []
for f in range(folds):
    xs = [rng.randn(n, d).astype(np.float32) for d in dims]
    c = np.stack([rng.uniform(22, 36, n).round(0), rng.randint(1, 3, n)], 1).astype(np.float32)
    fi = (rng.randn(n) * 15 + 105).astype(np.float32)
    ct = torch.from_numpy(c).to(dev)
    specs.append(MemberSpec(dims, [110, 110], 10, 2, [pack_rows(torch.from_numpy(x).to(dev), ct) for x in xs], combine="gpoe",
                            batch=b, seed=f, head="regression", y=torch.from_numpy(fi).to(dev),
                            row_order=torch.from_numpy(loader_orders(n, b, 3 * epochs, 3)).to(dev)))
out = {"folds": folds, "epochs": epochs, "rows": n, "batch": b, "dims": dims, "steps_per_epoch": -(-n // b)}
for eng, flag in (("tcgen05-generic", _lib.TRAIN_TC_SIMPLE), ("fp32", _lib.TRAIN_FP32)):
    tr = EnsembleTrainer(specs, device=dev)
    tr.train_epochs(epochs, flags=flag)           # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.train_epochs(epochs, flags=flag); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = epochs * out["steps_per_epoch"]
    out[eng] = {"ms": ms, "us_per_fold_step": 1e3 * ms / steps, "samples_per_s": folds * epochs * n / (ms / 1e3)}
    tr.close()
# the end-to-end supervised model (dual decoders + BatchNorm / dropout classifier + contrastive hinge), K folds, one launch
specs = []
for f in range(folds):
    xs = [rng.randn(n, d).astype(np.float32) for d in dims]
    c = np.zeros((n, 29), np.float32); c[np.arange(n), rng.randint(0, 27, n)] = 1; c[np.arange(n), 27 + rng.randint(0, 2, n)] = 1
    ct = torch.from_numpy(c).to(dev)
    specs.append(MemberSpec(dims, [110, 110], 10, 29, [pack_rows(torch.from_numpy(x).to(dev), ct) for x in xs], combine="poe",
                            batch=256, seed=f, head="endtoend", head_hidden=[128, 64, 32],
                            head_params=dict(margin=1.0, w_contrastive=1.0, w_kl=0.1, w_rec=0.1, dropout=0.5),
                            y=torch.from_numpy((rng.rand(n) > 0.5).astype(np.float32)).to(dev)))
spe2 = -(-n // 256)
for eng, flag in (("endtoend fp32", _lib.TRAIN_FP32), ("endtoend tcgen05-generic", _lib.TRAIN_TC_SIMPLE)):
    tr = EnsembleTrainer(specs, device=dev)
    tr.train_epochs(epochs, flags=flag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.train_epochs(epochs, flags=flag); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[eng] = {"ms": ms, "batch": 256, "us_per_fold_step": 1e3 * ms / (epochs * spe2), "samples_per_s": folds * epochs * n / (ms / 1e3)}
    tr.close()
from oracle import cvae_torch
model = cvae_torch.OracleCVAEEndToEnd(dims, [110, 110], 10, 29, 1e-4, 3, non_linear=True, classifier_layers=[128, 64, 32])
model.train()
xs = [torch.from_numpy(rng.randn(n, d).astype(np.float32)) for d in dims]
c = torch.zeros(n, 29); c[:, 0] = 1
lab = torch.from_numpy((rng.rand(n) > 0.5).astype(np.int64))
eps = rng.randn(3 * spe2, 256, 10).astype(np.float32); keep = (rng.rand(3 * spe2, 256, 224) > 0.5).astype(np.float32)
t0 = time.perf_counter()
cvae_torch.e2e_train_loop(model, xs, c, lab, 256, 3, eps, keep, [128, 64, 32], 1.0, 1.0)
dt = time.perf_counter() - t0
out["endtoend cpu_oracle_torch"] = {"threads": torch.get_num_threads(), "us_per_fold_step": 1e6 * dt / (3 * spe2), "samples_per_s": 3 * n / dt}
# CPU: the oracle's restatement of the reference loop, one fold, all host threads
from oracle import cvae_torch
model = cvae_torch.OracleCVAERegression(dims, [110, 110], 10, 2, 1e-4, 3, non_linear=True)
xs = [torch.from_numpy(rng.randn(n, d).astype(np.float32)) for d in dims]
c = torch.from_numpy(np.stack([rng.uniform(22, 36, n).round(0), rng.randint(1, 3, n)], 1).astype(np.float32))
fi = torch.from_numpy((rng.randn(n) * 15 + 105).astype(np.float32))
order = loader_orders(n, b, 3, 3)
eps = rng.randn(3 * out["steps_per_epoch"], b, 10).astype(np.float32)
t0 = time.perf_counter()
cvae_torch.regression_train_loop(model, xs, c, fi, order, "gpoe", b, eps)
dt = time.perf_counter() - t0
out["cpu_oracle_torch"] = {"threads": torch.get_num_threads(), "us_per_fold_step": 1e6 * dt / (3 * out["steps_per_epoch"]),
                           "samples_per_s": 3 * n / dt}
print(json.dumps(out, indent=1))
