"""f3: pin the oracle's restatement of ``cVAE_multimodal_regression`` (oracle/cvae_torch.py::OracleCVAERegression,
regression_train_loop) against vectors recorded from the UNMODIFIED reference class driven like
multimodal_kfold_train_cvae_supervised_regression.py:119-152 (oracle/make_golden.py --f3)."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, sub
from oracle import cvae_torch

CASES = ["reg_M3_full_gpoe", "reg_M2_small_poe"]


def build(g):
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    return cvae_torch.OracleCVAERegression(dims, [int(h) for h in g["hidden"]], int(g["z"]), 2, 1e-4, len(dims),
                                           non_linear=True), dims


@pytest.mark.parametrize("name", CASES)
def test_regression_oracle_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    model, dims = build(g)
    init = sub(g, "init/")
    sd = model.state_dict()
    assert set(sd) == set(init)                       # same parameter names as the reference's state_dict
    for k, v in init.items():
        assert np.array_equal(sd[k].numpy(), v), k    # seed-exact constructor (RNG order encoders, decoders, alphas, regressor)
    xs = [torch.from_numpy(g[f"x{i}"]) for i in range(len(dims))]
    c, fi = torch.from_numpy(g["c"]), torch.from_numpy(g["fi"])
    comb, b = str(g["combine"]), int(g["batch"])
    # step 0: losses + every gradient
    idx = [torch.from_numpy(g["order"][0, m, :b].astype(np.int64)) for m in range(len(dims))]
    out = model.step_losses([xs[m][idx[m]] for m in range(len(dims))], [c[idx[m]] for m in range(len(dims))], comb,
                            torch.from_numpy(g["eps"][0][: len(idx[0])]), fi[idx[0]])
    np.testing.assert_allclose([float(out["total"]), float(out["kl"]), float(out["ll"]), float(out["regression"])],
                               g["losses"][0], rtol=1e-5)
    np.testing.assert_allclose(out["fi_pred"].detach().numpy(), g["fi_pred0"], rtol=1e-4, atol=1e-5)
    model.optimizer1.zero_grad()
    out["total"].backward()
    params = dict(model.named_parameters())
    for k, v in sub(g, "grad/").items():
        np.testing.assert_allclose(params[k].grad.numpy(), v, rtol=2e-4, atol=2e-6 * np.abs(v).max(), err_msg=k)
    model.optimizer1.zero_grad()
    log = cvae_torch.regression_train_loop(model, xs, c, fi, g["order"], comb, b, g["eps"])
    np.testing.assert_allclose(log, g["losses"], rtol=2e-5)
    g0 = sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, model.state_dict()[k].numpy(), v, init[k], len(log), 1e-4, False, g0.get(k))
    # evaluation pass (eval mode still samples z, ..._regression.py:146-152)
    model.eval()
    with torch.no_grad():
        xt = [torch.from_numpy(g[f"xt{i}"]) for i in range(len(dims))]
        ev = model.step_losses(xt, [torch.from_numpy(g["ct"])] * len(dims), comb, torch.from_numpy(g["eps_test"]))
    np.testing.assert_allclose(ev["fi_pred"].numpy(), g["fi_pred_test"], rtol=1e-4, atol=1e-4)
    for i in range(len(dims)):
        np.testing.assert_allclose(ev["x_recons"][i].numpy(), g[f"pred{i}"], rtol=1e-4, atol=1e-5)
