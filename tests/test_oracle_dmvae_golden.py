"""f4: pin the oracle's restatement of the DMVAE family (oracle/cvae_torch.py::OracleDMVAE) against vectors recorded from
the UNMODIFIED ``DMVAE`` / ``mmVAEPlus`` / ``WeightedDMVAE`` classes (oracle/make_golden.py --f4b)."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, sub
from oracle import cvae_torch

CASES = ["dmvae_M2_shared", "mmvaeplus_M3_shared", "wdmvae_M3_shared", "dmvae_M3_default"]


def build(g):
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    return cvae_torch.OracleDMVAE(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["s_dim"]), 1e-4, len(dims),
                                  beta=float(g["beta"]), weighted=str(g["cls"]) == "WeightedDMVAE"), dims


@pytest.mark.parametrize("name", CASES)
def test_dmvae_family_oracle_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    model, dims = build(g)
    init = sub(g, "init/")
    sd = model.state_dict()
    assert set(sd) == set(init)
    for k, v in init.items():
        assert np.array_equal(sd[k].numpy(), v), k           # seed-exact: encoders, decoders, then WeightedDMVAE.weights
    z, s_dim = int(g["z"]), int(g["s_dim"])
    zc = max(0, z - s_dim)
    if name == "dmvae_M3_default":
        assert zc == 0 and float(np.abs(g["losses"][:, 1]).max()) == 0.0       # no shared latent: M deterministic autoencoders
    xs = [torch.from_numpy(g[f"x{i}"]) for i in range(len(dims))]
    b = int(g["batch"])
    out = model.step_losses([x[:b] for x in xs], torch.from_numpy(g["eps"][0][:b, :zc].copy()))
    np.testing.assert_allclose([float(out["total"].detach()), float(torch.as_tensor(out["kl"]).detach()), float(out["ll"].detach())],
                               g["losses"][0], rtol=1e-5, atol=1e-7)
    model.optimizer1.zero_grad()
    out["total"].backward()
    for k, p in model.named_parameters():
        want = g["grad/" + k]
        if "nograd/" + k in g:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            np.testing.assert_allclose(p.grad.numpy(), want, rtol=2e-4, atol=2e-6 * np.abs(want).max(), err_msg=k)
    model.optimizer1.zero_grad()
    log = cvae_torch.dmvae_train_loop(model, xs, b, int(g["epochs"]), g["eps"], zc)
    np.testing.assert_allclose(log, g["losses"], rtol=2e-5, atol=1e-7)
    g0 = sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, model.state_dict()[k].numpy(), v, init[k], len(log), 1e-4, False, g0.get(k))
