"""Dev tool: BASELINE configs[4] ("scale stress": 100 000 subjects x 1 000 features, 5 folds) timed on one GPU -- 5 folds x S
seeds members, each with its own 80 000 training rows (313 minibatches per epoch) and 20 000 test rows: one epoch of the
pipelined training kernel, then reconstruction + deviation + AUC of every member's test rows."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, scoring, workloads

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(5)
n_tr, n_te, d, c_dim, z, folds = 80000, 20000, 1000, 29, 10, 5

def rows(n):
    x = torch.randn(n, d, device=dev, generator=g)
    c = torch.zeros(n, c_dim, device=dev)
    c[torch.arange(n, device=dev), torch.randint(0, 27, (n,), device=dev, generator=g)] = 1
    c[torch.arange(n, device=dev), 27 + torch.randint(0, 2, (n,), device=dev, generator=g)] = 1
    return pack_rows(x, c)
train = [rows(n_tr) for _ in range(folds)]
test = [rows(n_te) for _ in range(folds)]
labels = [(torch.rand(n_te, device=dev, generator=g) < 0.3).to(torch.uint8) for _ in range(folds)]
specs = [MemberSpec([d], [110, 110], z, c_dim, [train[f]], seed=100 * f + s, state_dict=workloads.init_state_dict(d, (110, 110), z, c_dim, 42 + s))
         for f in range(folds) for s in range(seeds)]
tr = EnsembleTrainer(specs, device=dev)
assert tr.engine() == "tcgen05-pipelined"
tr.train_steps(8); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tr.train_epochs(1); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
test_xc = [[test[f]] for f in range(folds) for _ in range(seeds)]
def score():
    xh, _, _ = tr.reconstruct(test_xc, mode="sample")
    roi, _, subj = scoring.deviation([t[0] for t in test_xc], [h[0] for h in xh])
    return scoring.auc(subj, [labels[f] for f in range(folds) for _ in range(seeds)])
score(); torch.cuda.synchronize()
e0.record(); auc = score(); e1.record(); torch.cuda.synchronize()
ms_s = e0.elapsed_time(e1)
flops = len(specs) * n_tr * workloads.train_flops_per_sample(d, c_dim, (110, 110), z)
byts = len(specs) * workloads.train_bytes_per_epoch(n_tr, 256, d, c_dim, (110, 110), z)
print(json.dumps({"config": "cfg5: 100 000 subjects x 1 000 features, 5 folds", "members": len(specs), "seeds_per_fold": seeds,
                  "train_rows_per_member": n_tr, "steps_per_epoch": 313, "epoch_ms": ms, "train_samples_per_s": len(specs) * n_tr / (ms * 1e-3),
                  "us_per_member_step": 1e3 * ms / 313, "algorithmic_TFLOPs": flops / (ms * 1e-3) / 1e12,
                  "algorithmic_GBps": byts / (ms * 1e-3) / 1e9, "scoring_ms": ms_s, "scored_subjects_per_s": len(specs) * n_te / (ms_s * 1e-3),
                  "device_memory_GB": torch.cuda.max_memory_allocated(dev) / 1e9}, indent=1))
