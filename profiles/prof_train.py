"""Small driver for ncu: N ensemble members of the cfg4 mix, a few fused train launches, one
reconstruction + deviation pass.  Usage: python profiles/prof_train.py [seeds] [minibatch steps] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_modal_normative_modeling_b200 import EnsembleTrainer, scoring, workloads

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 15
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=seeds)
tr = EnsembleTrainer(wl.specs, device=dev)
for _ in range(launches):
    tr.train_steps(steps)
torch.cuda.synchronize()
train_xc = [s.xc for s in wl.specs]
xhat_tr, _, _ = tr.reconstruct(train_xc, mode="mean")
xhat_te, _, _ = tr.reconstruct(wl.test_xc, mode="mean")
stats = scoring.normative_stats([t[0] for t in train_xc], [h[0] for h in xhat_tr], wl.train_hc_mask)
roi, z, subj = scoring.deviation([t[0] for t in wl.test_xc], [h[0] for h in xhat_te], stats)
auc = scoring.auc(z, wl.test_labels)
torch.cuda.synchronize()
print("members", tr.n, "ok", float(torch.cat([a.mean()[None] for a in auc]).mean()))
tr.close()
