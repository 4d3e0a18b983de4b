"""Dev tool: train the cfg4 workload with each engine from identical state / Philox draws and
compare per-member losses and final parameters.  Usage: python tools/cmp_engines.py [seeds] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=seeds)
res = {}
for name, flags in (("fp32", _lib.TRAIN_FP32), ("tcp", 0), ("tcs", _lib.TRAIN_TC_SIMPLE)):
    tr = EnsembleTrainer(wl.specs, device=dev)
    print(name, tr.engine(flags))
    losses = tr.train_steps(steps, record_losses=True, flags=flags)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train_steps(steps, flags=flags)
    torch.cuda.synchronize()
    print(name, "second call ms", 1e3 * (time.perf_counter() - t0))
    res[name] = (losses.cpu().numpy(), tr.params.cpu().numpy().copy())
    tr.close()
ref_l, ref_p = res["fp32"]
for name in ("tcp", "tcs"):
    l, p = res[name]
    rel = np.abs(l[:, :, 0] - ref_l[:, :, 0]) / np.abs(ref_l[:, :, 0])
    worst = np.argsort(-rel.max(axis=1))[:8]
    print(name, "max rel loss err", rel.max(), "worst members", [(int(i), wl.tags[i][1][:5], float(rel[i].max())) for i in worst])
    print(name, "per-step max rel err", rel.max(axis=0))
    print(name, "param max abs diff", np.abs(p - ref_p).max())
