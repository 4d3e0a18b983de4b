"""Dev tool: one step (no Adam), gradients of each engine vs the FP32 engine, per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib

dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=1)
idx = [i for i, t in enumerate(wl.tags) if t[0] == 0]          # fold 0: 3 x D=116 + early fusion
specs = [wl.specs[i] for i in idx]
F = _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS
out = {}
for name, flags in (("fp32", _lib.TRAIN_FP32), ("tcp", 0), ("tcs", _lib.TRAIN_TC_SIMPLE)):
    tr = EnsembleTrainer(specs, device=dev, keep_grads=True)
    for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
        tr.train_steps(1, flags=flags | F)
    torch.cuda.synchronize()
    out[name] = [tr.state_dict(i, "grads") for i in range(len(specs))]
    tr.close()
for name in ("tcp", "tcs"):
    for i in range(len(specs)):
        for k, v in out["fp32"][i].items():
            a = out[name][i][k].cpu().numpy().astype(np.float64); b = v.cpu().numpy().astype(np.float64)
            err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
            if err > 2e-5:
                bad = np.argwhere(np.abs(a - b) > 1e-4 * np.abs(b).max())
                print(name, i, wl.tags[idx[i]][1][:6], k, "relerr %.2e" % err, "n_bad", len(bad), "first", bad[:3].tolist(),
                      "last", bad[-2:].tolist())
print("done")
