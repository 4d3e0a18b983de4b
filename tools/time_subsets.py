"""Dev tool: time the pipelined kernel on subsets of the cfg4 ensemble (per-architecture cost at full chip)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib
if os.environ.get('NMB_LIB'): _lib.LIB_PATH = os.path.abspath(os.environ['NMB_LIB'])   # A/B against another build

dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=24)
steps = 20
def run(idx, label):
    specs = [wl.specs[i] for i in idx]
    tr = EnsembleTrainer(specs, device=dev)
    for _ in range(3):
        tr.train_steps(steps)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        tr.train_steps(steps)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    n = len(specs)
    waves = -(-n // 148)
    print("%-28s members %4d  ms/launch %7.3f  us per member-step per SM %7.1f  (waves %d -> %.1f us/step/wave)" %
          (label, n, ms, ms * 1e3 * min(n, 148) / (n * steps), waves, ms * 1e3 / (waves * steps)))
    tr.close()
small = [i for i, t in enumerate(wl.tags) if not t[1].startswith("early")]
big = [i for i, t in enumerate(wl.tags) if t[1].startswith("early")]
quick = "quick" in sys.argv
run(small[:148], "D=116 x148 (1 wave)")
if not quick:
    run(small[:296], "D=116 x296 (2 waves)")
run(small[:37], "D=116 x37 (quarter chip)")
run(big[:120], "D=348 x120")
if not quick:
    run(big[:30], "D=348 x30")
run(list(range(len(wl.specs))), "cfg4 all 480")
