import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib

def run(dims, n, b, steps, comb="poe", c_dim=29, seed=0):
    rng = np.random.RandomState(seed)
    c = np.zeros((n, c_dim), np.float32); c[np.arange(n), rng.randint(0, c_dim - 2, n)] = 1; c[np.arange(n), c_dim - 2 + rng.randint(0, 2, n)] = 1
    ct = torch.from_numpy(c).cuda()
    xc = [pack_rows(torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda(), ct) for d in dims]
    eps = torch.from_numpy(rng.randn(1, steps, b, 10).astype(np.float32)).cuda()
    out = {}
    for eng, fl in (("tcs", _lib.TRAIN_TC_SIMPLE), ("fp32", _lib.TRAIN_FP32)):
        torch.manual_seed(1)
        tr = EnsembleTrainer([MemberSpec(dims, [110, 110], 10, c_dim, xc, combine=comb, batch=b, seed=1)])
        tr.params.copy_(torch.from_numpy(np.random.RandomState(5).randn(tr.total_params).astype(np.float32) * 0.05).cuda())
        tr.train_steps(steps, eps=eps, flags=fl)
        out[eng] = {k: v.cpu().numpy() for k, v in tr.state_dict(0).items()}
        tr.close()
    res = []
    for k in out["tcs"]:
        if k.endswith("encoder_layers.0.weight") or k.endswith("decoder_mean_layer.weight"):
            d = np.abs(out["tcs"][k] - out["fp32"][k]) / (steps * 1e-4)
            res.append("%s max %.1e n>1e-2 %d" % (k.replace("encoder_list", "enc").replace("decoder_list", "dec").replace("encoder_layers", "l").replace("decoder_mean_layer", "out"), d.max(), int((d > 1e-2).sum())))
    print(dims, "n", n, "b", b, "steps", steps, comb, "|", "; ".join(res))

run([116, 58, 30], 150, 128, 2)
run([116, 58, 30], 256, 128, 2)
run([116, 58, 30], 160, 128, 2)
run([116, 58, 30], 150, 128, 1)
run([116], 150, 128, 2)
run([58, 116, 30], 150, 128, 2)
run([116, 116, 116], 300, 256, 2)
run([116, 116, 116], 300, 256, 2, "gpoe")
run([116, 116, 116], 512, 256, 2)
