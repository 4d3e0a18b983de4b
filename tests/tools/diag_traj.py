import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import load, sub
from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "mmjsd_M3"
comb = sys.argv[2] if len(sys.argv) > 2 else "poe"
g = load(os.path.join(R, "tests/golden"), name)
dims = [int(d) for d in g["dims"]]
c = torch.from_numpy(g["c"]).cuda()
xc = [pack_rows(torch.from_numpy(g[f"x{i}"]).cuda(), c) for i in range(len(dims))]
sd = {k: torch.from_numpy(v) for k, v in sub(g, "init/").items()}
def mk(keep=False):
    return EnsembleTrainer([MemberSpec(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), xc, combine=comb,
                                       batch=int(g["batch"]), seed=1, state_dict=sd)], keep_grads=keep)
# gradient accuracy by magnitude class, step 0
ref = sub(g, "grad/")
for eng, fl in (("tc", 0), ("tcs", _lib.TRAIN_TC_SIMPLE), ("fp32", _lib.TRAIN_FP32)):
    tr = mk(True)
    tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], flags=fl | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS)
    gr = tr.state_dict(0, "grads")
    for k in ("encoder_list.0.encoder_layers.0.weight", "encoder_list.1.encoder_layers.0.weight", "encoder_list.0.encoder_layers.1.weight",
              "decoder_list.0.decoder_layers.0.weight"):
        v = ref[k]; got = gr[k].cpu().numpy().reshape(v.shape)
        rel = np.abs(got - v) / (np.abs(v) + 1e-30)
        mx = np.abs(v).max()
        small = np.abs(v) < 1e-2 * mx
        print(eng, k, "max-norm rel %.1e | elementwise rel: median %.1e q99 %.1e | frac |g|<1e-2 max: %.3f, their median rel %.1e, sign flips %d" % (
            np.abs(got - v).max() / mx, np.median(rel), np.quantile(rel, .99), small.mean(), np.median(rel[small]) if small.any() else 0,
            int((np.sign(got) != np.sign(v)).sum())))
    tr.close()
    for steps in (1, g["eps"].shape[0]):
        tr = mk()
        tr.train_steps(steps, eps=torch.from_numpy(g["eps"][:steps]).cuda()[None], flags=fl)
        sdo, init = tr.state_dict(0), sub(g, "init/")
        if steps == g["eps"].shape[0]:
            for k in ("encoder_list.0.encoder_layers.0.weight", "encoder_list.1.encoder_layers.0.weight", "encoder_list.2.encoder_layers.0.weight",
                      "encoder_list.0.encoder_layers.1.weight", "decoder_list.0.decoder_layers.0.weight", "decoder_list.0.decoder_mean_layer.weight"):
                v = g["final/" + k]
                d_ref, d_got = v - init[k], sdo[k].cpu().numpy().reshape(v.shape) - init[k]
                dev = np.sort((np.abs(d_got - d_ref) / np.abs(d_ref).max()).ravel())
                print("  ", eng, steps, "%-45s q50 %.1e q90 %.1e q99 %.1e max %.1e" % (k, dev[dev.size // 2], dev[int(dev.size * .9)], dev[int(dev.size * .99)], dev[-1]))
        tr.close()
print("---- pattern of the deviation in encoder_list.0.encoder_layers.0.weight (tcs)")
for steps in (1, 2, 3, 4):
    tr = mk()
    tr.train_steps(steps, eps=torch.from_numpy(g["eps"][:steps]).cuda()[None], flags=_lib.TRAIN_TC_SIMPLE)
    trf = mk()
    trf.train_steps(steps, eps=torch.from_numpy(g["eps"][:steps]).cuda()[None], flags=_lib.TRAIN_FP32)
    for k in ("encoder_list.0.encoder_layers.0.weight", "encoder_list.1.encoder_layers.0.weight"):
        a, b = tr.state_dict(0)[k].cpu().numpy(), trf.state_dict(0)[k].cpu().numpy()
        d = np.abs(a - b) / (steps * 1e-4)
        print(steps, k, "vs fp32 engine: max %.2e; worst columns" % d.max(), np.argsort(-d.max(0))[:8], "col max", np.sort(d.max(0))[-8:].round(3),
              "worst rows", np.argsort(-d.max(1))[:5], "n elems > 1e-2:", int((d > 1e-2).sum()))
    tr.close(); trf.close()
print("---- gradient of step 2 (ragged) at the weights after one FP32-engine Adam step: tcs vs fp32 engine")
gs = {}
for eng, fl in (("tcs", _lib.TRAIN_TC_SIMPLE), ("fp32", _lib.TRAIN_FP32)):
    tr = mk(True)
    tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], flags=_lib.TRAIN_FP32)
    tr.grads.zero_()
    tr.train_steps(1, eps=torch.from_numpy(g["eps"][1:2]).cuda()[None], flags=fl | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS)
    gs[eng] = {k: v.cpu().numpy() for k, v in tr.state_dict(0, "grads").items()}
    m1 = {k: v.cpu().numpy() for k, v in tr.state_dict(0, "adam_m").items()}
    tr.close()
for k in ("encoder_list.0.encoder_layers.0.weight", "encoder_list.1.encoder_layers.0.weight", "encoder_list.0.encoder_layers.1.weight"):
    a, b = gs["tcs"][k], gs["fp32"][k]
    mx = np.abs(b).max()
    rel = np.abs(a - b) / mx
    # Adam after step 2: m = .9 m1 + .1 g2 with m1 = .1 g1: the update direction is sensitive where g2 ~ -.9 g1... look at it
    mm = 0.9 * m1[k] + 0.1 * b
    print(k, "grad max-norm rel err: max %.1e median %.1e | |g2|/max: median %.2e | frac of elements with |m2| < 1e-2 max|m2|: %.3f" % (
        rel.max(), np.median(rel), np.median(np.abs(b)) / mx, (np.abs(mm) < 1e-2 * np.abs(mm).max()).mean()))
