"""f4: pin the oracle's restatement of ``mvtCAE`` (oracle/cvae_torch.py::OracleMvtCAE) against vectors recorded from the
UNMODIFIED class (oracle/make_golden.py --f4c): regular fusion (gPoE, MoPoE), the 'poe' branch as written (clamped with
two experts' worth of precision or more, unclamped with one) and the total-correlation term."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, sub
from oracle import cvae_torch

CASES = ["mvtcae_M3_gpoe", "mvtcae_M2_poe", "mvtcae_M3_mopoe", "mvtcae_M1_poe"]


@pytest.mark.parametrize("name", CASES)
def test_mvtcae_oracle_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    model = cvae_torch.OracleMvtCAE(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), 1e-4, len(dims), non_linear=True)
    model.optimizer1 = torch.optim.Adam(model.parameters(), lr=1e-4)          # mvtCAE: Adam(self.parameters()) (:1774)
    assert np.array_equal(torch.randn(4).numpy(), g["next_draw"])
    init = sub(g, "init/")
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k
    xs = [torch.from_numpy(g[f"x{i}"]) for i in range(len(dims))]
    cs = [torch.from_numpy(g["c"]).long() for _ in dims]
    comb, b, n = str(g["combine"]), int(g["batch"]), int(g["n"])
    log, s = [], 0
    for _ in range(int(g["epochs"])):
        for r0 in range(0, n, b):
            rows = min(b, n - r0)
            out = model.step_losses([x[r0:r0 + rows] for x in xs], [c[r0:r0 + rows] for c in cs], comb, torch.from_numpy(g["eps"][s][:rows]))
            model.optimizer1.zero_grad()
            out["total"].backward()
            if s == 0:
                np.testing.assert_allclose(out["logvar"].detach().numpy(), g["logvar"], rtol=1e-4, atol=1e-5)
                for k, p in model.named_parameters():
                    v = g["grad/" + k]
                    got = p.grad.numpy() if p.grad is not None else np.zeros_like(v)
                    np.testing.assert_allclose(got, v, rtol=3e-4, atol=3e-6 * max(np.abs(v).max(), 1e-12), err_msg=k)
            model.optimizer1.step()
            log.append([float(out[k].detach()) for k in ("total", "kl", "ll", "tc")])
            s += 1
    np.testing.assert_allclose(np.asarray(log), g["losses"], rtol=3e-5)
    g0 = sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, model.state_dict()[k].numpy(), v, init[k], len(log), 1e-4, False, g0.get(k))
