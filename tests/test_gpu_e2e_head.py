"""f3: the end-to-end supervised model -- ``cVAE_multimodal_endtoend`` v2 + ``Classifier`` (cVAE.py:2004-2207)
trained like multimodal_kfold_cvae_nmpmcont.py:226-247 and evaluated with ``predict`` (:30-46) -- through the C ABI on both
generic engines, against vectors recorded from the unmodified reference (oracle/make_golden.py --f3e): dual decoder
sets, BatchNorm (batch statistics, running statistics), injected dropout masks, cross-entropy, contrastive hinge."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, relerr, sub
from test_oracle_e2e_golden import is_dead_bias

pytestmark = pytest.mark.gpu
REL = 1e-4
CASES = ["e2e_M3_full", "e2e_M2_small"]
ENGINES = ["tcs", "fp32"]
COLS = (0, 1, 3, 4, 5, 6)        # LOSS8 columns holding (total, kl, ce, rec_health, rec_disease, contrastive)


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
                      combine="poe", batch=int(g["batch"]), seed=5, state_dict=sd, head="endtoend",
                      head_hidden=[int(w) for w in g["layers"]],
                      head_params=dict(margin=float(g["margin"]), w_contrastive=float(g["w_con"]), w_kl=0.1, w_rec=0.1,
                                       dropout=float(g["dropout"])),
                      y=torch.from_numpy(g["labels"].astype(np.float32)).cuda(), drop_keep=torch.from_numpy(g["keep"]).cuda())
    return EnsembleTrainer([spec], keep_grads=keep_grads), xc


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_e2e_step_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr, _ = make_trainer(g)
    assert tr.engine() == "tcgen05-generic"
    flags = (engine_flags(engine) | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS | _lib.TRAIN_LOSS8
             | _lib.TRAIN_NO_STATS)
    losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    got = losses[0, 0].cpu().numpy()
    assert np.allclose(got[list(COLS)], g["losses"][0], rtol=REL), (got, g["losses"][0])
    assert np.isclose(got[2], -(g["losses"][0][3] + g["losses"][0][4]), rtol=REL)
    mu, lv, xr = tr.peek(0)
    m = len(g["dims"])
    assert relerr(mu.cpu().numpy(), g["mu"]) < REL and relerr(lv.cpu().numpy(), g["logvar"]) < REL
    for i in range(m):
        assert relerr(xr[i].cpu().numpy(), g[f"xh_health{i}"]) < REL
        assert relerr(xr[m + i].cpu().numpy(), g[f"xh_disease{i}"]) < REL
    assert relerr(tr.peek_head(0).cpu().numpy(), g["logits0"]) < 2 * REL
    grads = tr.state_dict(0, "grads")
    ref = sub(g, "grad/")
    for k, v in ref.items():
        if is_dead_bias(k, len(g["layers"])):
            assert np.abs(grads[k].cpu().numpy()).max() < 1e-4 * max(np.abs(ref[k[:-4] + "weight"]).max(), 1e-6)
            continue
        got_k = grads[k].cpu().numpy().reshape(v.shape)
        assert np.abs(got_k - v).max() / (np.abs(v).max() + 1e-30) < 2 * REL, (k, np.abs(got_k - v).max() / np.abs(v).max())
    sd = tr.state_dict(0)
    for k, v in sub(g, "init/").items():          # NO_ADAM + NO_STATS: nothing moved, running statistics included
        assert np.array_equal(sd[k].cpu().numpy().reshape(v.shape), v.astype(np.float32)), k
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_e2e_epochs_with_adam_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr, _ = make_trainer(g, keep_grads=False)
    steps = g["eps"].shape[0]
    losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"]).cuda()[None], record_losses=True,
                            flags=engine_flags(engine) | _lib.TRAIN_LOSS8)
    torch.cuda.synchronize()
    got, want = losses[0].cpu().numpy().astype(np.float64)[:, list(COLS)], g["losses"]
    for col, rel in ((0, REL), (3, REL), (4, REL), (1, 10 * REL), (2, 10 * REL), (5, 10 * REL)):
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, init, g0 = tr.state_dict(0), sub(g, "init/"), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        got_k = sd[k].cpu().numpy().reshape(v.shape)
        if k.endswith("num_batches_tracked"):
            assert int(got_k) == int(v) == steps
        elif "running_" in k:
            # running mean: carries the dead bias of its Linear, which random-walks by +-lr per step in BOTH implementations
            assert np.allclose(got_k, v, rtol=2e-4, atol=steps * 1e-4 if "mean" in k else 1e-6), k
        elif not is_dead_bias(k, len(g["layers"])):
            # FP32 engine: 99 % of the elements within 2e-4 of the largest update (measured: 1e-5).  BF16x3 engine: losses
            # agree to 1e-6 and the decoders to 1e-5, but units that come within rounding distance of a ReLU / leaky-relu
            # kink after the first Adam steps take the other branch (the fixture's knife-edge scan covers the initial
            # weights only); in the 44-row ragged batch one sample is 1/44 of the gradient, and Adam's normalisation turns
            # the rank-one difference into visible steps on the encoders (measured q99 3e-2 on encoder 0, 4e-3 elsewhere;
            # same mechanism as tests/tools/diag_traj.py shows for the mmJSD fixture).  The end-to-end program trains on the
            # bit-stable FP32 engine (e2e.py); the BF16x3 engine is held to the looser bound below.
            assert_update_close(k, got_k, v, init[k], steps, 1e-4, engine == "fp32", g0.get(k), q99_tc=6e-2, mean_tc=1e-2)
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_e2e_predict_vs_reference(golden_dir, name, engine):
    """``predict`` in eval mode (running statistics, no dropout, classifier on the fused mean) and the two decoder sets of
    an eval-mode forward."""
    from multi_modal_normative_modeling_b200 import pack_rows
    g = load(golden_dir, name)
    tr, _ = make_trainer(g, sd_prefix="final/", keep_grads=False)
    dims = [int(d) for d in g["dims"]]
    m = len(dims)
    ct = torch.from_numpy(g["ct"]).cuda()
    xt = [pack_rows(torch.from_numpy(g[f"xt{i}"]).cuda(), ct) for i in range(m)]
    logits = tr.head_predict([xt], mode="mean", engine=engine)
    torch.cuda.synchronize()
    assert logits[0].shape == (g["logits_test"].shape[0], 2)
    assert np.abs(logits[0].cpu().numpy() - g["logits_test"]).max() < 1e-3 * (np.abs(g["logits_test"]).max() + 1)
    assert (logits[0].argmax(1).cpu().numpy() == g["logits_test"].argmax(1)).mean() > 0.97
    lg, xh = tr.head_predict([xt], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()], engine=engine, want_xhat=True)
    for i in range(m):
        assert relerr(xh[0][i].cpu().numpy(), g[f"pred_health{i}"]) < REL
        assert relerr(xh[0][m + i].cpu().numpy(), g[f"pred_disease{i}"]) < REL
    assert np.abs(lg[0].cpu().numpy() - g["logits_test_sampled"]).max() < 1e-3 * (np.abs(g["logits_test_sampled"]).max() + 1)
    tr.close()


def test_e2e_philox_dropout_rate():
    """Without injected masks the kernel draws the dropout keep flags from Philox stream 2: the trained model still
    predicts, training is reproducible (same seed, same trajectory), and a different seed gives a different one."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.randn(64, 12).astype(np.float32)).cuda()
    c = torch.zeros(64, 3).cuda(); c[:, 0] = 1
    y = torch.from_numpy((rng.rand(64) > 0.5).astype(np.float32)).cuda()
    xc = [pack_rows(x, c)]

    def run(seed):
        torch.manual_seed(1)
        sd = None
        tr = EnsembleTrainer([MemberSpec([12], [9], 3, 3, xc, batch=32, seed=seed, head="endtoend", head_hidden=[16, 8], y=y)])
        tr.params.copy_(torch.from_numpy(np.random.RandomState(3).randn(tr.total_params).astype(np.float32) * 0.1).cuda())
        v = tr._views(0, tr.params)
        for k in v:
            if k.endswith("running_var") or (k.endswith(".weight") and k.count(".") == 3 and k.split(".")[2] in ("1", "5")):
                v[k].fill_(1.0)
        losses = tr.train_steps(6, record_losses=True, flags=_lib.TRAIN_TC_SIMPLE | _lib.TRAIN_LOSS8)
        torch.cuda.synchronize()
        out = losses[0].cpu().numpy().copy()
        tr.close()
        return out
    a, b, c2 = run(7), run(7), run(8)
    assert np.isfinite(a).all() and np.array_equal(a, b) and not np.array_equal(a, c2)


def test_dropin_e2e_module_loop_vs_reference(golden_dir):
    """Drop-in class through the reference's loop body (nmpmcont :226-247): forward -> loss_function(xs, fwd, labels, margin,
    weightcontrastive) -> zero_grad -> backward -> optimizer.step(), with the recorded eps / dropout draws fed through
    torch.randn / torch.rand; then evaluate()'s predict under model.eval() and an eval-mode forward."""
    import cVAE as shim                                    # the reference's module name exports the v2 class
    name = "e2e_M2_small"
    g = load(golden_dir, name)
    dims, widths = [int(d) for d in g["dims"]], [int(w) for w in g["layers"]]
    p = float(g["dropout"])
    torch.manual_seed(int(g["seed"]))
    model = shim.cVAE_multimodal_endtoend(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), modalities=len(dims),
                                          non_linear=True, classifier_layers=widths, dropout_rate=p, num_classes=2)
    init = sub(g, "init/")
    assert set(model.state_dict()) == set(init)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k
    model.to("cuda").train()
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    c, lab = torch.from_numpy(g["c"]).cuda(), torch.from_numpy(g["labels"]).cuda()
    n, b = int(g["n"]), int(g["batch"])
    real_randn, real_rand = torch.randn, torch.rand
    log, s = [], 0
    try:
        for _ in range(int(g["epochs"])):
            for r0 in range(0, n, b):
                rows = min(b, n - r0)
                torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows]).to(k.get("device", "cpu"))
                # keep = (rand >= p): feed u = 1 where the reference kept the unit, 0 where it dropped it
                torch.rand = lambda *a, **k: torch.from_numpy(g["keep"][s][:rows]).to(k.get("device", "cpu"))
                xb, cb = [x[r0:r0 + rows] for x in xs], [c[r0:r0 + rows]] * len(dims)
                fwd = model.forward(xb, cb)
                torch.randn, torch.rand = real_randn, real_rand
                losses = model.loss_function(xb, fwd, lab[r0:r0 + rows], float(g["margin"]), float(g["w_con"]))
                if s == 0:
                    assert relerr(fwd["logits"].cpu().numpy(), g["logits0"]) < 2 * REL
                    assert relerr(fwd["mu"].cpu().numpy(), g["mu"]) < REL
                    assert relerr(fwd["x_recons_disease"][1].loc.cpu().numpy(), g["xh_disease1"]) < REL
                    with pytest.raises(ValueError, match="LAST forward"):
                        model.loss_function(xb, dict(fwd), lab[r0:r0 + rows])
                model.optimizer.zero_grad()
                losses["total_loss"].backward()
                if s == 0:
                    for k, pr in model.named_parameters():
                        v = g["grad/" + k]
                        if not is_dead_bias(k, len(widths)):
                            assert np.abs(pr.grad.cpu().numpy() - v).max() / (np.abs(v).max() + 1e-30) < 2 * REL, k
                model.optimizer.step()
                log.append([float(losses[k].detach()) for k in ("total_loss", "kl_loss", "classification_loss", "recon_loss_health",
                                                               "recon_loss_disease", "contrastive_loss")])
                s += 1
    finally:
        torch.randn, torch.rand = real_randn, real_rand
    got, want = np.asarray(log), g["losses"]
    for col, rel in ((0, REL), (3, REL), (4, REL), (1, 10 * REL), (2, 10 * REL), (5, 10 * REL)):
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, g0 = model.state_dict(), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        got_k = sd[k].cpu().numpy()
        if k.endswith("num_batches_tracked"):
            assert int(got_k) == int(v) == s                   # one running-statistics update per training step
        elif "running_" in k:
            assert np.allclose(got_k, v, rtol=2e-4, atol=s * 1e-4 if "mean" in k else 1e-6), k
        elif not is_dead_bias(k, len(widths)):
            assert_update_close(k, got_k, v, init[k], s, 1e-4, True, g0.get(k))        # the drop-in trains on the FP32 engine
    model.eval()
    ct = torch.from_numpy(g["ct"]).cuda()
    xt = [torch.from_numpy(g[f"xt{i}"]).cuda() for i in range(len(dims))]
    logits = model.predict(xt, [ct] * len(dims))
    assert np.abs(logits.cpu().numpy() - g["logits_test"]).max() < 1e-3 * (np.abs(g["logits_test"]).max() + 1)
    torch.randn = lambda *a, **k: torch.from_numpy(g["eps_test"]).to(k.get("device", "cpu"))
    try:
        ev = model.forward(xt, [ct] * len(dims))
    finally:
        torch.randn = real_randn
    assert relerr(ev["x_recons_health"][0].loc.cpu().numpy(), g["pred_health0"]) < REL
    assert np.abs(ev["logits"].cpu().numpy() - g["logits_test_sampled"]).max() < 1e-3 * (np.abs(g["logits_test_sampled"]).max() + 1)
    model.close()


def test_e2e_program_end_to_end_vs_oracle(tmp_path):
    """multimodal_kfold_cvae_nmpmcont.py on a synthetic HCPimage-shaped dataset: every fold in one launch, production
    Philox eps (stream 0) and dropout (stream 2).  Fold 0 is replayed through the oracle with the documented streams."""
    from multi_modal_normative_modeling_b200 import e2e, synthetic
    from oracle import cvae_torch, philox
    synthetic.write_dataset(str(tmp_path), "HCPimage", n=160, seed=5)
    args = e2e.build_parser().parse_args(["-R", "HCPimage", "-P", "SE-PoE", "-E", "2", "-K", "2", "-H", "110", "110", "10",
                                          "-Layers", "16", "8", "-Margin", "0.7", "-Weightcontrastive", "0.5"])
    dbg = {}
    metrics = e2e.main(args, root=tmp_path, debug=dbg)
    assert len(metrics) == 2 and all(0.0 <= m["accuracy"] <= 1.0 for m in metrics)
    assert (tmp_path / "results_endtoend.csv").read_text().count("accuracy $") == 1
    assert (tmp_path / "outputs" / "kfold_analysis_endtoend" / "train_ids_001.csv").exists()
    fd = dbg["folds"][0]
    dims, z, b, widths, p = fd["dims"], 10, 256, [16, 8], 0.5
    model = cvae_torch.OracleCVAEEndToEnd(dims, [110, 110], z, 29, 1e-4, len(dims), non_linear=True, classifier_layers=widths,
                                          dropout_rate=p)
    model.load_state_dict(fd["init"])
    model.train()
    xs = [torch.from_numpy(x[:, :d].copy()) for x, d in zip(fd["xc_train"], dims)]
    c = torch.from_numpy(fd["xc_train"][0][:, dims[0]:dims[0] + 29].copy())
    n = xs[0].shape[0]
    spe = -(-n // b)
    eps = np.stack([philox.normals(fd["seed"], s, b * z, 0).reshape(b, z) for s in range(2 * spe)])
    keep = np.stack([(philox.uniforms(fd["seed"], s, b * sum(widths), 2) >= p).astype(np.float32).reshape(b, sum(widths))
                     for s in range(2 * spe)])
    log = cvae_torch.e2e_train_loop(model, xs, c, torch.from_numpy(fd["labels"]), b, 2, eps, keep, widths, 0.7, 0.5)
    got = dbg["losses"][0][: 2 * spe].astype(np.float64)[:, list(COLS)]
    assert np.allclose(got[:, 0], log[:, 0], rtol=2e-4), (got, log)
    assert np.allclose(got[:, 2], log[:, 2], rtol=2e-3) and np.allclose(got[:, 5], log[:, 5], rtol=2e-3), (got, log)
    model.eval()
    xt = [torch.from_numpy(x[:, :d].copy()) for x, d in zip(fd["xc_test"], dims)]
    ct = torch.from_numpy(fd["xc_test"][0][:, dims[0]:dims[0] + 29].copy())
    want = model.predict(xt, [ct] * len(dims)).numpy()
    assert np.abs(dbg["logits"][0] - want).max() < 2e-3 * (np.abs(want).max() + 1.0)
    # the saved module is the drop-in class with the trained weights: it predicts what the ensemble predicted
    saved = torch.load(tmp_path / "outputs" / "kfold_analysis" / "supervised_cvae" / "000" / "cVAE_model.pkl", weights_only=False)
    saved.to("cuda").eval()
    again = saved.predict([t.cuda() for t in xt], [ct.cuda()] * len(dims)).cpu().numpy()
    assert np.abs(again - dbg["logits"][0]).max() < 1e-5 * (np.abs(want).max() + 1.0)
