"""Dev tool: N Adam steps per engine from identical state; per-tensor parameter drift vs FP32 engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib

dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=1)
idx = [i for i, t in enumerate(wl.tags) if t[0] == 0]
specs = [wl.specs[i] for i in idx]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out = {}
for name, flags in (("fp32", _lib.TRAIN_FP32), ("tcp", 0), ("tcs", _lib.TRAIN_TC_SIMPLE)):
    tr = EnsembleTrainer(specs, device=dev)
    tr.train_steps(n, flags=flags)
    torch.cuda.synchronize()
    out[name] = [tr.state_dict(i) for i in range(len(specs))]
    tr.close()
for i in (0, 3):
    for k, v in out["fp32"][i].items():
        b = v.cpu().numpy()
        d1 = np.abs(out["tcp"][i][k].cpu().numpy() - b); d2 = np.abs(out["tcs"][i][k].cpu().numpy() - b)
        print(i, k, "tcp max %.2e mean %.2e n>1e-4 %d | tcs max %.2e mean %.2e n>1e-4 %d" %
              (d1.max(), d1.mean(), (d1 > 1e-4).sum(), d2.max(), d2.mean(), (d2 > 1e-4).sum()))
