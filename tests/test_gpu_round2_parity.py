"""GPU parity at FULL size for every BASELINE config shape, against vectors recorded from the unmodified
reference (oracle/make_golden.py::round2_cases), for all three engines, through the C ABI:

  cfg1 D=150, cfg2 / cfg4 early fusion D=348, cfg5 D=1000, cfg3 four modalities 116/116/116/348 with gPoE / PoE /
  MoE / MoPoE, latent 32 (unfused latent items), 4 hidden layers, hidden > 127 (generic engine), the nmmlp class
  (-MSE), each with batch 256 (or 128 / 64) AND a ragged last batch.

Tolerance: 1e-4 relative per-step forward / backward (north_star); gradient rows of knife-edge leaky-relu units
(recorded in the fixture, see helpers.drop_knife_rows) are skipped.
"""
import numpy as np
import pytest
import torch

from helpers import (GENERIC_ONLY, LOOP_CASES, assert_grads_close, assert_losses_close, assert_update_close, is_nmmlp,
                     load, relerr, sub)

pytestmark = pytest.mark.gpu
REL = 1e-4
ENGINES = ["tc", "tcs", "fp32"]


def engine_flags(engine):
    from multi_modal_normative_modeling_b200 import _lib
    return {"tc": 0, "tcs": _lib.TRAIN_TC_SIMPLE, "fp32": _lib.TRAIN_FP32}[engine]


def make_trainer(g, name, sd_prefix="init/", keep_grads=True):
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    dims = [int(d) for d in g["dims"]]
    c = torch.from_numpy(g["c"]).cuda()
    xc = [pack_rows(torch.from_numpy(g[f"x{i}"]).cuda(), c) for i in range(len(dims))]
    sd = {k: torch.from_numpy(v) for k, v in sub(g, sd_prefix).items()}
    spec = MemberSpec(input_dims=dims, hidden=[int(h) for h in g["hidden"]], latent=int(g["z"]), c_dim=int(g["c_dim"]),
                      xc=xc, combine=str(g["combine"]), loss_kind="neg_mse" if is_nmmlp(name) else "gauss_ll",
                      batch=int(g["batch"]), seed=11, state_dict=sd)
    return EnsembleTrainer([spec], keep_grads=keep_grads), xc


def check_grads(g, prefix, grads):
    assert_grads_close(g, prefix, grads, REL, to_numpy=lambda t: t.cpu().numpy())


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", LOOP_CASES)
def test_step_full_and_ragged_batch_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr, _ = make_trainer(g, name)
    if engine == "tc":
        assert tr.engine() == ("tcgen05-generic" if name in GENERIC_ONLY else "tcgen05-pipelined")
    flags = engine_flags(engine) | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS
    b = int(g["batch"])
    # step 0: the first full batch on the initial weights
    losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    assert np.allclose(losses[0, 0].cpu().numpy(), g["losses"][0], rtol=REL), (losses[0, 0], g["losses"][0])
    mu, lv, xr = tr.peek(0)
    assert mu.shape[0] == b
    assert relerr(mu.cpu().numpy(), g["mu"]) < REL
    assert relerr(lv.cpu().numpy(), g["logvar"]) < REL
    for i in range(len(xr)):
        if f"xrecon{i}" in g:
            assert relerr(xr[i].cpu().numpy(), g[f"xrecon{i}"]) < REL
    check_grads(g, "grad/", tr.state_dict(0, "grads"))
    # ... then every later step of the epoch up to the ragged last batch, still on the initial weights (NO_ADAM)
    rs = int(g["ragged_step"])
    for s in range(1, rs + 1):
        tr.grads.zero_()
        losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][s:s + 1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    assert np.allclose(losses[0, 0].cpu().numpy(), g["lossr"], rtol=REL), (losses[0, 0], g["lossr"])
    _, _, xr = tr.peek(0)
    assert xr[0].shape[0] == int(g["n"]) - rs * b          # the partial batch
    if sub(g, "gradr/"):
        check_grads(g, "gradr/", tr.state_dict(0, "grads"))
    for k, v in sub(g, "init/").items():                   # NO_ADAM leaves the parameters bit-identical
        assert np.array_equal(tr.state_dict(0)[k].cpu().numpy().reshape(v.shape), v), k
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", LOOP_CASES)
def test_epochs_with_adam_vs_reference(golden_dir, name, engine):
    """Two epochs of the reference loop (full + ragged batches, Adam): per-step losses and the parameter UPDATE against
    the reference's recording; bounds and their derivation in helpers.assert_losses_close / assert_update_close."""
    g = load(golden_dir, name)
    tr, _ = make_trainer(g, name, keep_grads=False)
    steps = g["eps"].shape[0]
    losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"]).cuda()[None], record_losses=True,
                            flags=engine_flags(engine))
    torch.cuda.synchronize()
    assert_losses_close(losses[0].cpu().numpy(), g["losses"], REL)
    assert int(tr.steps_done()[0]) == steps
    final = sub(g, "final/")
    if not final:
        tr.close()
        return
    sd, init, g0 = tr.state_dict(0), sub(g, "init/"), sub(g, "grad/")
    for k, v in final.items():
        got = sd[k].cpu().numpy().reshape(v.shape)
        assert_update_close(k, got, v, init[k], steps, 1e-4, engine == "fp32", g0.get(k))
    tr.close()


@pytest.mark.parametrize("name", [c for c in LOOP_CASES if c not in ("mm_M4_full_poe", "mm_M4_full_moe", "mm_M4_full_mopoe")])
def test_reconstruct_and_deviation_vs_reference(golden_dir, name):
    from multi_modal_normative_modeling_b200 import scoring
    g = load(golden_dir, name)
    tr, xc = make_trainer(g, name, sd_prefix="final/", keep_grads=False)
    xhat, mu, lv = tr.reconstruct([xc], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()], want_latent=True)
    torch.cuda.synchronize()
    for i in range(len(xc)):
        assert relerr(xhat[0][i].cpu().numpy(), g[f"pred{i}"]) < REL, i
    _, _, subj = scoring.deviation(xc, xhat[0], want_roi=False)
    for i in range(len(xc)):
        assert relerr(subj[i].cpu().numpy(), g[f"dev{i}"]) < 5e-4
    ref_hat = [torch.from_numpy(g[f"pred{i}"]).cuda() for i in range(len(xc))]
    roi, _, subj = scoring.deviation(xc, ref_hat)
    for i in range(len(xc)):
        assert relerr(subj[i].cpu().numpy(), g[f"dev{i}"]) < 1e-5
        if f"dev_roi{i}" in g:
            assert relerr(roi[i].cpu().numpy(), g[f"dev_roi{i}"]) < 1e-5
    tr.close()


@pytest.mark.parametrize("name", ["cvae_D116_full", "cvae_D150_full"])
def test_dropin_cvae_forward_loss_backward_step(golden_dir, name):
    """Drop-in signature #1: cVAE.forward -> loss_function -> zero_grad -> backward ->
    optimizer1.step() on the product class against the reference's recording of the same calls (cVAE.py:435-504)."""
    from multi_modal_normative_modeling_b200.cVAE import cVAE
    g = load(golden_dir, name)
    torch.manual_seed(int(g["seed"]))
    model = cVAE(int(g["d"]), [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), learning_rate=1e-4,
                 non_linear=True)
    assert np.array_equal(torch.randn(4).numpy(), g["next_draw"])          # same generator state as the reference
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), g["init/" + k]), k
    model.to("cuda")
    x, c = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["c"]).long().cuda()
    real_randn = torch.randn
    torch.randn = lambda *a, **k: torch.from_numpy(g["eps"]).to(k.get("device", "cpu"))
    try:
        fwd = model.forward(x, c)
    finally:
        torch.randn = real_randn
    loss = model.loss_function(x, fwd)
    assert set(loss) == {"total", "kl", "ll"} and tuple(loss["total"].shape) == (1,) and tuple(loss["ll"].shape) == (1,)
    assert np.allclose([float(loss["total"]), float(loss["kl"]), float(loss["ll"])], g["losses"], rtol=REL)
    assert relerr(fwd["mu"].detach().cpu().numpy(), g["mu"]) < REL
    assert relerr(fwd["logvar"].detach().cpu().numpy(), g["logvar"]) < REL
    assert relerr(fwd["x_recon"].loc.detach().cpu().numpy(), g["xrecon"]) < REL
    assert relerr(fwd["x_recon"].scale.detach().cpu().numpy(), np.exp(g["init/decoder.logvar_out"]) ** 0.5) < 1e-6
    model.optimizer1.zero_grad()
    loss["total"].backward()
    for k, p in model.named_parameters():
        if "grad/" + k in g:
            assert relerr(p.grad.cpu().numpy(), g["grad/" + k]) < REL, k
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k          # discriminator: untouched
    model.optimizer1.step()
    for k, v in model.state_dict().items():
        assert_update_close(k, v.cpu().numpy(), g["final/" + k], g["init/" + k], 1, 1e-4, False, g.get("grad/" + k))
    # a stale fwd_rtn is refused (loss_function must describe the LAST forward)
    with pytest.raises(ValueError):
        model.loss_function(x, {"x_recon": fwd["x_recon"], "mu": fwd["mu"].clone(), "logvar": fwd["logvar"]})
    # encode / reparameterise / decode pieces (cVAE.py:415-428)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "final/").items()})
    mu, logvar = model.encode(x, c)
    assert relerr(mu.detach().cpu().numpy(), g["latent"]) < REL
    assert relerr(logvar.detach().exp().cpu().numpy(), g["latent_var"]) < REL
    rec = model.decode(mu, c)
    assert relerr(rec.loc.detach().cpu().numpy(), g["pred"]) < REL
    torch.manual_seed(3)
    e = torch.randn_like(mu)
    torch.manual_seed(3)
    z = model.reparameterise(mu, logvar)
    assert torch.allclose(z, mu + e * torch.exp(0.5 * logvar), rtol=1e-6, atol=1e-7)
