"""Workload of tools/sanitize.sh: a 160-member cfg4-shaped ensemble (more members than SMs -> chunked work items with
the cross-SM acquire / release hand-off), a few minibatch steps on the pipelined kernel, then one scoring pass
(pipelined forward-only reconstruct, normative statistics, deviation / z, AUC)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, scoring, workloads

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=seeds)
tr = EnsembleTrainer(wl.specs, device=dev)
losses = tr.train_steps(steps, record_losses=True)
sc = scoring.DeviationScorer(tr, [s.xc for s in wl.specs], wl.test_xc, wl.train_hc_mask, wl.test_labels, mode="sample").run()
torch.cuda.synchronize()
print("ok", tr.engine(), tr.n, "members", steps, "steps; mean loss", float(losses[:, -1, 0].mean()),
      "mean subject AUC", float(sc.auc_subj.mean()))
