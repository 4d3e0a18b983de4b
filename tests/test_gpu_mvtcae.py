"""f4: the ``mvtCAE`` baseline (cVAE.py:1754-1893) through the C ABI (``NMB_FAMILY_MVTCAE``) on both generic
engines and through the drop-in class, against vectors recorded from the unmodified reference class
(oracle/make_golden.py --f4c): clamped fused variance, the 'poe' branch as written, +1e-5 * ll, the tc term."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, relerr, sub

pytestmark = pytest.mark.gpu
REL = 1e-4
CASES = ["mvtcae_M3_gpoe", "mvtcae_M2_poe", "mvtcae_M3_mopoe", "mvtcae_M1_poe"]
ENGINES = ["tcs", "fp32"]


def engine_flags(engine):
    from multi_modal_normative_modeling_b200 import _lib
    return {"tcs": _lib.TRAIN_TC_SIMPLE, "fp32": _lib.TRAIN_FP32}[engine]


def make_trainer(g, sd_prefix="init/", keep_grads=True):
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    dims = [int(d) for d in g["dims"]]
    c = torch.from_numpy(g["c"]).cuda()
    xc = [pack_rows(torch.from_numpy(g[f"x{i}"]).cuda(), c) for i in range(len(dims))]
    sd = {k: torch.from_numpy(v) for k, v in sub(g, sd_prefix).items()}
    spec = MemberSpec(input_dims=dims, hidden=[int(h) for h in g["hidden"]], latent=int(g["z"]), c_dim=int(g["c_dim"]), xc=xc,
                      combine=str(g["combine"]), batch=int(g["batch"]), seed=3, state_dict=sd, family="mvtcae", beta=float(g["beta"]))
    return EnsembleTrainer([spec], keep_grads=keep_grads), xc


def check_grads(g, grads):
    ref = sub(g, "grad/")
    alphas = sorted(k for k in ref if k.startswith("alpha_m_list."))
    want = np.concatenate([ref[k].ravel() for k in alphas]); got = np.concatenate([grads[k].cpu().numpy().ravel() for k in alphas])
    assert np.abs(got - want).max() <= REL * np.abs(want).max() + 1e-12, "alpha_m_list"
    for k, v in ref.items():
        if k in alphas:
            continue
        gk = grads[k].cpu().numpy().reshape(v.shape)
        assert np.abs(gk - v).max() / (np.abs(v).max() + 1e-30) < 2 * REL, (k, np.abs(gk - v).max() / np.abs(v).max())


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_mvtcae_step_epochs_prediction_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr, _ = make_trainer(g)
    assert tr.engine() == "tcgen05-generic"
    flags = engine_flags(engine) | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS | _lib.TRAIN_LOSS4
    losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    assert np.allclose(losses[0, 0].cpu().numpy(), g["losses"][0], rtol=REL), (losses[0, 0], g["losses"][0])    # total, kl, ll, tc
    mu, lv, _ = tr.peek(0)
    assert relerr(mu.cpu().numpy(), g["mu"]) < REL and relerr(lv.cpu().numpy(), g["logvar"]) < REL
    check_grads(g, tr.state_dict(0, "grads"))
    tr.close()
    tr, _ = make_trainer(g, keep_grads=False)
    steps = g["eps"].shape[0]
    losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"]).cuda()[None], record_losses=True, flags=engine_flags(engine) | _lib.TRAIN_LOSS4)
    got, want = losses[0].cpu().numpy().astype(np.float64), g["losses"]
    for col, rel in ((0, 10 * REL), (1, 10 * REL), (2, REL), (3, 10 * REL)):
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, init, g0 = tr.state_dict(0), sub(g, "init/"), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, sd[k].cpu().numpy().reshape(v.shape), v, init[k], steps, 1e-4, engine == "fp32", g0.get(k),
                            q99_tc=6e-2, mean_tc=1e-2)
    tr.close()
    tr, xc = make_trainer(g, sd_prefix="final/", keep_grads=False)
    xhat, _, _ = tr.reconstruct([xc], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()], engine="fp32" if engine == "fp32" else "tcs")
    for i in range(len(xc)):
        assert relerr(xhat[0][i].cpu().numpy(), g[f"pred{i}"]) < 2 * REL, i
    tr.close()


@pytest.mark.parametrize("name", ["mvtcae_M3_gpoe", "mvtcae_M2_poe"])
def test_mvtcae_dropin_vs_reference(golden_dir, name):
    import cVAE as shim
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    model = shim.mvtCAE(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), learning_rate=1e-4, modalities=len(dims),
                        non_linear=True)
    init = sub(g, "init/")
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k
    model.to("cuda")
    model._engine_flags = _lib.TRAIN_FP32
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    c = torch.from_numpy(g["c"]).long().cuda()
    n, b, comb = int(g["n"]), int(g["batch"]), str(g["combine"])
    real = torch.randn
    log, s = [], 0
    try:
        for _ in range(int(g["epochs"])):
            for r0 in range(0, n, b):
                rows = min(b, n - r0)
                torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows]).to(k.get("device", "cpu"))
                fwd = model.forward_multimodal([x[r0:r0 + rows] for x in xs], [c[r0:r0 + rows]] * len(dims), comb)
                torch.randn = real
                loss = model.loss_function_multimodal([x[r0:r0 + rows] for x in xs], fwd)
                model.optimizer1.zero_grad()
                loss["total"].backward()
                model.optimizer1.step()
                log.append([float(loss[k].detach()) for k in ("total", "kl", "ll", "tc")])
                s += 1
    finally:
        torch.randn = real
    got, want = np.asarray(log), g["losses"]
    for col, rel in ((0, 10 * REL), (1, 10 * REL), (2, REL), (3, 10 * REL)):
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, g0 = model.state_dict(), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, sd[k].cpu().numpy(), v, init[k], s, 1e-4, True, g0.get(k))
    model.close()
