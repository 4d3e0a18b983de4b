"""Dev tool: scale-stress shape (BASELINE.json configs[4], reduced): D = 1000 features, many rows, a few members;
pipelined engine vs FP32 engine on losses, and throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
members = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d, c_dim = 1000, 29
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda()
c = torch.zeros(n, c_dim).cuda(); c[torch.arange(n), torch.from_numpy(rng.randint(0, 27, n)).cuda()] = 1; c[:, 27] = 1
xc = [pack_rows(x, c)]
sd = workloads.init_state_dict(d, (110, 110), 10, c_dim, 42)
specs = [MemberSpec([d], [110, 110], 10, c_dim, xc, batch=256, seed=k, state_dict=sd) for k in range(members)]
res = {}
for name, flags in (("fp32", _lib.TRAIN_FP32), ("tcp", 0)):
    tr = EnsembleTrainer(specs)
    print(name, tr.engine(flags))
    steps = 24
    losses = tr.train_steps(steps, record_losses=True, flags=flags)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.train_steps(steps, flags=flags); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "ms per launch", dt * 1e3, "samples/s", members * steps * 256 / dt)
    res[name] = losses.cpu().numpy()
    tr.close()
rel = np.abs(res["tcp"][:, :, 0] - res["fp32"][:, :, 0]) / np.abs(res["fp32"][:, :, 0])
print("max rel loss diff", rel.max(), "finite", np.isfinite(res["tcp"]).all())
