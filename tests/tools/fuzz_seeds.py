"""Dev tool: one fuzz configuration under many data seeds; prints the margins of the test's criteria."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import cvae_torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
dims, hidden, z, c_dim, combine, loss, n, batch = [129], [64], 16, 3, "poe", "gauss_ll", 257, 256
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    rng = np.random.RandomState(seed)
    c = np.zeros((n, c_dim), np.float32); c[np.arange(n), rng.randint(0, c_dim, n)] = 1
    cc = torch.from_numpy(c).cuda()
    xc = [pack_rows(torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda(), cc) for d in dims]
    spe = -(-n // batch); steps = 2 * spe + 1
    specs = []
    for k in range(3):
        torch.manual_seed(11 + k)
        model = cvae_torch.OracleCVAEMultimodal(dims, hidden, z, c_dim, 1e-4, len(dims), True, loss)
        sd = {a: b.detach().clone() for a, b in model.state_dict().items()}
        specs.append(MemberSpec(dims, hidden, z, c_dim, xc, combine=combine, loss_kind=loss, batch=batch, seed=5 + k, state_dict=sd))
    out = {}
    for name, flags in (("fp32", _lib.TRAIN_FP32), ("tc", 0), ("tc2", 0), ("tcs", _lib.TRAIN_TC_SIMPLE)):
        tr = EnsembleTrainer(specs)
        a = tr.train_steps(steps, record_losses=True, flags=flags); torch.cuda.synchronize()
        out[name] = (a.cpu().numpy(), tr.params.cpu().numpy().copy()); tr.close()
    l0, p0 = out["fp32"]
    for name in ("tc", "tcs"):
        l1, p1 = out[name]
        rel = np.abs(l1 - l0) / (np.abs(l0) + 1e-30)
        d = np.abs(p1 - p0)
        print(seed, name, "loss rel max %.2e" % rel.max(), "param d max %.2e q995 %.2e" % (d.max(), np.quantile(d, 0.995)),
              "deterministic" if name != "tc" else ("det=%s" % np.array_equal(out["tc"][1], out["tc2"][1])))
