"""f3: the supervised regression head -- ``cVAE_multimodal_regression`` (cVAE.py:2211-2347) trained like
multimodal_kfold_train_cvae_supervised_regression.py:119-131 (per-modality shuffling loaders, target from modality 0)
and evaluated like :137-152 -- through the C ABI on both generic engines, against vectors recorded from the unmodified
reference (oracle/make_golden.py --f3).  Tolerance 1e-4 relative per step (north_star); Adam trajectories by
helpers.assert_update_close."""
import numpy as np
import pytest
import torch

from helpers import assert_grads_close, assert_update_close, load, relerr, sub

pytestmark = pytest.mark.gpu
REL = 1e-4
CASES = ["reg_M3_full_gpoe", "reg_M2_small_poe"]
ENGINES = ["tcs", "fp32"]


def engine_flags(engine):
    from multi_modal_normative_modeling_b200 import _lib
    return {"tcs": _lib.TRAIN_TC_SIMPLE, "fp32": _lib.TRAIN_FP32}[engine]


def make_trainer(g, sd_prefix="init/", keep_grads=True, order=True):
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    dims = [int(d) for d in g["dims"]]
    c = torch.from_numpy(g["c"]).cuda()
    xc = [pack_rows(torch.from_numpy(g[f"x{i}"]).cuda(), c) for i in range(len(dims))]
    sd = {k: torch.from_numpy(v) for k, v in sub(g, sd_prefix).items()}
    spec = MemberSpec(input_dims=dims, hidden=[int(h) for h in g["hidden"]], latent=int(g["z"]), c_dim=2, xc=xc,
                      combine=str(g["combine"]), batch=int(g["batch"]), seed=5, state_dict=sd, head="regression",
                      y=torch.from_numpy(g["fi"]).cuda(), row_order=torch.from_numpy(g["order"]).cuda() if order else None)
    return EnsembleTrainer([spec], keep_grads=keep_grads)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_regression_step_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr = make_trainer(g)
    assert tr.engine() == "tcgen05-generic"          # members with a head never take the pipelined kernel
    flags = engine_flags(engine) | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS | _lib.TRAIN_LOSS4
    losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    got = losses[0, 0].cpu().numpy()
    assert got.shape == (4,)
    assert np.allclose(got, g["losses"][0], rtol=REL), (got, g["losses"][0])      # total, kl, ll, regression
    mu, _, _ = tr.peek(0)
    assert relerr(mu.cpu().numpy(), g["mu"]) < REL
    grads = tr.state_dict(0, "grads")
    assert {k for k in grads if k.startswith("regressor.")} == {f"regressor.{l}.{w}" for l in (0, 2, 4) for w in ("weight", "bias")}
    assert_grads_close(g, "grad/", grads, REL, to_numpy=lambda t: t.cpu().numpy())
    for k, v in sub(g, "init/").items():
        assert np.array_equal(tr.state_dict(0)[k].cpu().numpy().reshape(v.shape), v), k
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_regression_epochs_with_adam_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr = make_trainer(g, keep_grads=False)
    steps = g["eps"].shape[0]
    losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"]).cuda()[None], record_losses=True,
                            flags=engine_flags(engine) | _lib.TRAIN_LOSS4)
    torch.cuda.synchronize()
    got, want = losses[0].cpu().numpy().astype(np.float64), g["losses"]
    for col, rel in ((0, REL), (2, REL), (3, 10 * REL), (1, 10 * REL)):      # kl / head loss after Adam steps: see helpers
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, init, g0 = tr.state_dict(0), sub(g, "init/"), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        # (BF16x3 engine: the head's MSE gradient (|fi_pred - fi| ~ 17 at initialisation) dominates every upstream
        #  gradient and changes sign-pattern from step to step, so a few more near-cancelling Adam averages flip than in
        #  the unsupervised cases: 99 % of the elements within 5e-3 instead of 2e-3 of the largest update)
        assert_update_close(k, sd[k].cpu().numpy().reshape(v.shape), v, init[k], steps, 1e-4, engine == "fp32", g0.get(k),
                            q99_tc=5e-3)
    # the per-epoch permutations cover exactly `epochs` epochs: one more step must be refused, not read out of bounds
    with pytest.raises(RuntimeError, match="row_order covers"):
        tr.train_steps(1, flags=engine_flags(engine))
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_regression_predict_vs_reference(golden_dir, name, engine):
    """Evaluation pass (..._regression.py:137-152): fi_pred and the reconstructions of the trained model."""
    from multi_modal_normative_modeling_b200 import pack_rows
    g = load(golden_dir, name)
    tr = make_trainer(g, sd_prefix="final/", keep_grads=False)
    dims = [int(d) for d in g["dims"]]
    ct = torch.from_numpy(g["ct"]).cuda()
    xt = [pack_rows(torch.from_numpy(g[f"xt{i}"]).cuda(), ct) for i in range(len(dims))]
    pred, xh = tr.head_predict([xt], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()], engine=engine, want_xhat=True)
    torch.cuda.synchronize()
    for i in range(len(dims)):
        assert relerr(xh[0][i].cpu().numpy(), g[f"pred{i}"]) < REL, i
    assert relerr(pred[0].cpu().numpy(), g["fi_pred_test"].ravel()) < 2 * REL
    only = tr.head_predict([xt], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()], engine=engine)
    assert torch.equal(only[0], pred[0])
    tr.close()


def test_head_members_are_validated():
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    x = torch.randn(20, 6, device="cuda"); c = torch.randn(20, 2, device="cuda")
    xc = [pack_rows(x, c)]
    with pytest.raises(ValueError, match="needs targets"):
        EnsembleTrainer([MemberSpec([6], [5], 3, 2, xc, head="regression")])
    with pytest.raises(RuntimeError, match="supervised head only"):
        EnsembleTrainer([MemberSpec([6], [5], 3, 2, xc, row_order=torch.zeros((1, 1, 20), dtype=torch.int32, device="cuda"))])
    tr = EnsembleTrainer([MemberSpec([6], [5], 3, 2, xc)])
    with pytest.raises(RuntimeError, match="LOSS4"):
        from multi_modal_normative_modeling_b200 import _lib
        tr.train_steps(1, record_losses=True, flags=_lib.TRAIN_LOSS4)
    tr.close()


def test_dropin_regression_module_loop_vs_reference(golden_dir):
    """Drop-in class through the reference's loop body (..._regression.py:119-131): forward_multimodal ->
    loss_function_multimodal(xes, out, fi, lambda_reg=1.0) -> zero_grad -> backward -> optimizer1.step(), eps drawn by
    torch (injected here through randn so that the recorded draws are used), against the reference's recording."""
    from multi_modal_normative_modeling_b200.cVAE import cVAE_multimodal_regression
    name = "reg_M2_small_poe"
    g = load(golden_dir, name)
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    model = cVAE_multimodal_regression(dims, [int(h) for h in g["hidden"]], int(g["z"]), 2, learning_rate=1e-4,
                                       modalities=len(dims), non_linear=True)
    init = sub(g, "init/")
    assert set(model.state_dict()) == set(init)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k           # seed-exact constructor
    model.to("cuda")
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    c, fi = torch.from_numpy(g["c"]).cuda(), torch.from_numpy(g["fi"]).cuda()
    n, b = int(g["n"]), int(g["batch"])
    real_randn = torch.randn
    s = 0
    log = []
    try:
        for ep in range(int(g["epochs"])):
            for r0 in range(0, n, b):
                idx = [torch.from_numpy(g["order"][ep, m, r0:r0 + b].astype(np.int64)).cuda() for m in range(len(dims))]
                rows = idx[0].numel()
                torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows]).to(k.get("device", "cpu"))
                out = model.forward_multimodal([xs[m][idx[m]] for m in range(len(dims))], [c[idx[m]] for m in range(len(dims))],
                                               str(g["combine"]))
                torch.randn = real_randn
                losses = model.loss_function_multimodal([xs[m][idx[m]] for m in range(len(dims))], out, fi[idx[0]], lambda_reg=1.0)
                if s == 0:
                    assert relerr(out["fi_pred"].cpu().numpy(), g["fi_pred0"]) < REL
                    assert relerr(out["mu_multimodal"].cpu().numpy(), g["mu"]) < REL
                    with pytest.raises(ValueError, match="LAST forward"):
                        model.loss_function_multimodal(xs, dict(out), fi[idx[0]])
                model.optimizer1.zero_grad()
                losses["total"].backward()
                if s == 0:
                    grads = {k: p.grad for k, p in model.named_parameters()}
                    assert_grads_close(g, "grad/", grads, REL, to_numpy=lambda t: t.cpu().numpy())
                model.optimizer1.step()
                log.append([float(losses[k]) for k in ("total", "kl", "ll", "regression")])
                s += 1
    finally:
        torch.randn = real_randn
    got, want = np.asarray(log), g["losses"]
    for col, rel in ((0, REL), (2, REL), (3, 10 * REL), (1, 10 * REL)):
        assert np.allclose(got[:, col], want[:, col], rtol=rel), (col, got[:, col], want[:, col])
    sd, g0 = model.state_dict(), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, sd[k].cpu().numpy(), v, init[k], s, 1e-4, False, g0.get(k), q99_tc=5e-3)
    model.close()


def test_regression_program_end_to_end_vs_oracle(tmp_path):
    """multimodal_kfold_train_cvae_supervised_regression.py on a synthetic HCPimage-shaped dataset: every fold in one
    launch, production Philox eps.  Fold 0 is replayed through the oracle (same initial weights, the same loader
    permutations, the documented Philox stream): per-step losses and test predictions."""
    from multi_modal_normative_modeling_b200 import regression, synthetic
    from oracle import cvae_torch, philox
    synthetic.write_dataset(str(tmp_path), "HCPimage", n=150, seed=3)
    args = regression.build_parser().parse_args(["-R", "HCPimage", "-P", "SE-gPoE", "-C", "gpoe", "-E", "2", "-K", "2",
                                                 "--batch_size", "64", "-H", "110", "110", "10"])
    dbg = {}
    scores = regression.train_and_test(args, root=tmp_path, debug=dbg)
    assert len(scores) == 2 and all(np.isfinite(list(s.values())).all() for s in scores)
    out = tmp_path / "regression_outputs"
    for f in range(2):
        pred, true = np.load(out / f"fold_{f}_pred.npy"), np.load(out / f"fold_{f}_true.npy")
        assert pred.shape == true.shape == (75, 1)
        for name in ("T1w_sMRI", "T2w_sMRI", "fMRI"):
            df = __import__("pandas").read_csv(out / f"deviation_fold_{f}_{name}_roiwise.csv")
            assert df.shape == (150, 117) and list(df.columns[:2]) == ["IID", "ROI_0"] and (df.iloc[:, 1:].to_numpy() >= 0).all()
    # ---- fold 0 through the oracle ----
    fd = dbg["folds"][0]
    dims, z, b = fd["dims"], 10, 64
    model = cvae_torch.OracleCVAERegression(dims, [110, 110], z, 2, 1e-4, len(dims), non_linear=True)
    model.load_state_dict(fd["init"])
    xs = [torch.from_numpy(x[:, :d].copy()) for x, d in zip(fd["xc_train"], dims)]
    c = torch.from_numpy(fd["xc_train"][0][:, dims[0]:dims[0] + 2].copy())
    n = xs[0].shape[0]
    spe = -(-n // b)
    eps = np.stack([philox.normals(fd["seed"], s, b * z, 0).reshape(b, z) for s in range(2 * spe)])
    log = cvae_torch.regression_train_loop(model, xs, c, torch.from_numpy(fd["fi"]), fd["order"], "gpoe", b, eps)
    got = dbg["losses"][0][: 2 * spe].astype(np.float64)
    assert np.allclose(got[:, 0], log[:, 0], rtol=5e-4), (got, log)
    assert np.allclose(got[:, 3], log[:, 3], rtol=2e-3), (got, log)
    xt = [torch.from_numpy(x[:, :d].copy()) for x, d in zip(fd["xc_test"], dims)]
    ct = torch.from_numpy(fd["xc_test"][0][:, dims[0]:dims[0] + 2].copy())
    eps_t = philox.normals(fd["seed"], 0, 256 * z, 1).reshape(256, z)[: xt[0].shape[0]]
    model.eval()
    with torch.no_grad():
        ev = model.step_losses(xt, [ct] * len(dims), "gpoe", torch.from_numpy(eps_t))
    want = ev["fi_pred"].numpy().ravel()
    assert np.abs(dbg["preds"][0] - want).max() < 2e-3 * (np.abs(want).max() + 1.0), (dbg["preds"][0][:5], want[:5])
