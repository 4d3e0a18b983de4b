"""f3: pin the oracle's restatement of ``cVAE_multimodal_endtoend`` v2 + ``Classifier`` (oracle/cvae_torch.py::
OracleCVAEEndToEnd, e2e_train_loop) against vectors recorded from the UNMODIFIED reference class driven like
multimodal_kfold_cvae_nmpmcont.py:226-247 and :30-46 (oracle/make_golden.py --f3e)."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, sub
from oracle import cvae_torch

CASES = ["e2e_M3_full", "e2e_M2_small"]


def is_dead_bias(k, n_hidden):
    """Linear biases in front of a BatchNorm: their gradient is exactly zero in exact arithmetic (the batch mean removes
    them), so the reference's value is rounding noise and Adam turns that noise into +-lr steps."""
    return any(k == f"classifier.classifier.{4 * l}.bias" for l in range(n_hidden))


def build(g):
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    return cvae_torch.OracleCVAEEndToEnd(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), 1e-4, len(dims),
                                         non_linear=True, classifier_layers=[int(w) for w in g["layers"]],
                                         dropout_rate=float(g["dropout"])), dims


@pytest.mark.parametrize("name", CASES)
def test_e2e_oracle_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    model, dims = build(g)
    init = sub(g, "init/")
    sd = model.state_dict()
    assert set(sd) == set(init)
    for k, v in init.items():
        assert np.array_equal(sd[k].numpy(), v), k            # seed-exact constructor
    widths = [int(w) for w in g["layers"]]
    xs = [torch.from_numpy(g[f"x{i}"]) for i in range(len(dims))]
    c, lab = torch.from_numpy(g["c"]), torch.from_numpy(g["labels"])
    b = int(g["batch"])
    model.train()
    keep, o = [], 0
    for w in widths:
        keep.append(torch.from_numpy(g["keep"][0][:b, o:o + w])); o += w
    state = {k: v.clone() for k, v in model.state_dict().items()}
    out = model.step_losses([x[:b] for x in xs], [c[:b]] * len(dims), lab[:b], torch.from_numpy(g["eps"][0][:b]), keep,
                            float(g["margin"]), float(g["w_con"]))
    got = [float(out[k].detach()) for k in ("total", "kl", "ce", "rec_health", "rec_disease", "contrastive")]
    np.testing.assert_allclose(got, g["losses"][0], rtol=1e-5)
    np.testing.assert_allclose(out["logits"].detach().numpy(), g["logits0"], rtol=1e-4, atol=1e-5)
    model.optimizer.zero_grad()
    out["total"].backward()
    params = dict(model.named_parameters())
    for k, v in sub(g, "grad/").items():
        if not is_dead_bias(k, len(widths)):
            np.testing.assert_allclose(params[k].grad.numpy(), v, rtol=2e-4, atol=2e-6 * np.abs(v).max(), err_msg=k)
    model.optimizer.zero_grad()
    model.load_state_dict(state)                              # undo the running-statistics update of the probe step
    log = cvae_torch.e2e_train_loop(model, xs, c, lab, b, int(g["epochs"]), g["eps"], g["keep"], widths, float(g["margin"]),
                                    float(g["w_con"]))
    np.testing.assert_allclose(log, g["losses"], rtol=5e-5)
    g0, sd = sub(g, "grad/"), model.state_dict()
    for k, v in sub(g, "final/").items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v)
        elif "running_" in k:
            # the running mean carries the dead bias of its Linear, which random-walks by +-lr per step (is_dead_bias)
            np.testing.assert_allclose(sd[k].numpy(), v, rtol=1e-4, atol=0.3 * len(log) * 1e-4 if "mean" in k else 1e-6)
        elif not is_dead_bias(k, len(widths)):
            assert_update_close(k, sd[k].numpy(), v, init[k], len(log), 1e-4, False, g0.get(k))
    model.eval()
    xt = [torch.from_numpy(g[f"xt{i}"]) for i in range(len(dims))]
    ct = torch.from_numpy(g["ct"])
    np.testing.assert_allclose(model.predict(xt, [ct] * len(dims)).numpy(), g["logits_test"], rtol=1e-3, atol=1e-4)
    with torch.no_grad():
        ev = model.step_losses(xt, [ct] * len(dims), eps=torch.from_numpy(g["eps_test"]))
    for i in range(len(dims)):
        np.testing.assert_allclose(ev["x_recons_health"][i].numpy(), g[f"pred_health{i}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(ev["x_recons_disease"][i].numpy(), g[f"pred_disease{i}"], rtol=1e-4, atol=1e-5)
