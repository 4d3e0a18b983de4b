"""Dev tool: run the pipelined training kernel a few times on the cfg4 workload (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib
if os.environ.get("NMB_LIB"): _lib.LIB_PATH = os.path.abspath(os.environ["NMB_LIB"])   # A/B against another build

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=seeds)
tr = EnsembleTrainer(wl.specs, device=dev)
for _ in range(reps):
    tr.train_steps(steps)
torch.cuda.synchronize()
print("ok", tr.engine())
