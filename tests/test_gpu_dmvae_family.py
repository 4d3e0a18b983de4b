"""f4: the DMVAE family of the baseline zoo -- ``DMVAE`` / ``mmVAEPlus`` / ``WeightedDMVAE`` (cVAE.py:1491-1752,
1895-2002) -- through the C ABI (``NMB_FAMILY_DMVAE``) on both generic engines, against vectors recorded from the
unmodified reference classes (oracle/make_golden.py --f4b): private / shared latent split, ProductOfExperts2,
sigmoid decoders with the 0.5 * SSE term, beta, learnable modality weights, and the default configuration whose
shared latent is empty."""
import numpy as np
import pytest
import torch

from helpers import assert_update_close, load, relerr, sub

pytestmark = pytest.mark.gpu
REL = 1e-4
CASES = ["dmvae_M2_shared", "mmvaeplus_M3_shared", "wdmvae_M3_shared", "dmvae_M3_default"]
ENGINES = ["tcs", "fp32"]


def engine_flags(engine):
    from multi_modal_normative_modeling_b200 import _lib
    return {"tcs": _lib.TRAIN_TC_SIMPLE, "fp32": _lib.TRAIN_FP32}[engine]


def make_trainer(g, sd_prefix="init/", keep_grads=True):
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    dims = [int(d) for d in g["dims"]]
    n = int(g["n"])
    none = torch.zeros((n, 0), device="cuda")
    xc = [pack_rows(torch.from_numpy(g[f"x{i}"]).cuda(), none) for i in range(len(dims))]
    sd = {k: torch.from_numpy(v) for k, v in sub(g, sd_prefix).items()}
    spec = MemberSpec(input_dims=dims, hidden=[int(h) for h in g["hidden"]], latent=int(g["z"]), c_dim=0, xc=xc,
                      batch=int(g["batch"]), seed=3, state_dict=sd, family="dmvae", s_dim=int(g["s_dim"]),
                      weighted=str(g["cls"]) == "WeightedDMVAE", beta=float(g["beta"]))
    return EnsembleTrainer([spec], keep_grads=keep_grads), xc


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_dmvae_family_step_vs_reference(golden_dir, name, engine):
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    tr, _ = make_trainer(g)
    assert tr.engine() == "tcgen05-generic"
    flags = engine_flags(engine) | _lib.TRAIN_NO_ADAM | _lib.TRAIN_WRITE_GRADS | _lib.TRAIN_KEEP_ACTS
    losses = tr.train_steps(1, eps=torch.from_numpy(g["eps"][:1]).cuda()[None], record_losses=True, flags=flags)
    torch.cuda.synchronize()
    assert np.allclose(losses[0, 0].cpu().numpy(), g["losses"][0], rtol=REL, atol=1e-7), (losses[0, 0], g["losses"][0])
    _, _, xr = tr.peek(0)
    for i in range(len(xr)):
        assert relerr(xr[i].cpu().numpy(), g[f"xrecon{i}"]) < REL
    grads = tr.state_dict(0, "grads")
    assert set(grads) == set(sub(g, "init/"))                          # the reference's state_dict names, nothing else
    for k, v in sub(g, "grad/").items():
        got = grads[k].cpu().numpy().reshape(v.shape)
        if np.abs(v).max() == 0:
            assert np.abs(got).max() == 0, k                           # e.g. fc_logvar when there is no shared latent
        else:
            assert np.abs(got - v).max() / np.abs(v).max() < REL, (k, np.abs(got - v).max() / np.abs(v).max())
    tr.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", CASES)
def test_dmvae_family_epochs_and_prediction_vs_reference(golden_dir, name, engine):
    g = load(golden_dir, name)
    tr, xc = make_trainer(g, keep_grads=False)
    steps = g["eps"].shape[0]
    losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"]).cuda()[None], record_losses=True, flags=engine_flags(engine))
    torch.cuda.synchronize()
    got, want = losses[0].cpu().numpy().astype(np.float64), g["losses"]
    assert np.allclose(got[:, 0], want[:, 0], rtol=REL) and np.allclose(got[:, 2], want[:, 2], rtol=REL), (got, want)
    assert np.allclose(got[:, 1], want[:, 1], rtol=10 * REL, atol=1e-7)
    sd, init, g0 = tr.state_dict(0), sub(g, "init/"), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, sd[k].cpu().numpy().reshape(v.shape), v, init[k], steps, 1e-4, engine == "fp32", g0.get(k),
                            q99_tc=6e-2, mean_tc=1e-2)     # BF16x3: ReLU units that come to the kink after the first steps
    tr.close()
    # pred_recon (:1574-1596) of the trained model: shared z sampled, private means
    tr, xc = make_trainer(g, sd_prefix="final/", keep_grads=False)
    xhat, _, _ = tr.reconstruct([xc], mode="sample", eps=[torch.from_numpy(g["eps_test"]).cuda()],
                                engine="fp32" if engine == "fp32" else "tcs")
    for i in range(len(xc)):
        assert relerr(xhat[0][i].cpu().numpy(), g[f"pred{i}"]) < REL, i
    tr.close()


@pytest.mark.parametrize("name", CASES)
def test_dmvae_family_dropins_vs_reference(golden_dir, name):
    """The drop-in classes through the reference loop (fwd -> loss -> zero_grad -> backward -> optimizer1.step()) with the
    recorded eps fed through torch.randn, then pred_recon; FP32 engine, so the trajectory is held to the exact bound."""
    import cVAE as shim
    import pandas as pd
    from multi_modal_normative_modeling_b200 import _lib
    g = load(golden_dir, name)
    dims = [int(d) for d in g["dims"]]
    z, s_dim = int(g["z"]), int(g["s_dim"])
    zc = max(0, z - s_dim)
    torch.manual_seed(int(g["seed"]))
    model = getattr(shim, str(g["cls"]))(dims, [int(h) for h in g["hidden"]], z, s_dim, learning_rate=1e-4, modalities=len(dims),
                                         non_linear=True)
    init = sub(g, "init/")
    assert set(model.state_dict()) == set(init)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k
    assert float(model.beta) == float(g["beta"])
    model.to("cuda")
    model._engine_flags = _lib.TRAIN_FP32
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    n, b = int(g["n"]), int(g["batch"])
    real = torch.randn
    log, s = [], 0
    try:
        for _ in range(int(g["epochs"])):
            for r0 in range(0, n, b):
                rows = min(b, n - r0)
                torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows, :zc].copy()).to(k.get("device", "cpu"))
                fwd = model.forward_multimodal([x[r0:r0 + rows] for x in xs], None, "poe")
                torch.randn = real
                loss = model.loss_function_multimodal([x[r0:r0 + rows] for x in xs], fwd)
                if s == 0:
                    assert fwd["mu_c"].shape == (rows, zc)
                    if zc:
                        assert relerr(fwd["mu_c"].detach().cpu().numpy(), g["mu_c"]) < REL
                    assert relerr(fwd["x_recons"][0].detach().cpu().numpy(), g["xrecon0"]) < REL
                model.optimizer1.zero_grad()
                loss["total"].backward()
                model.optimizer1.step()
                log.append([float(loss[k].detach()) for k in ("total", "kl", "ll")])
                s += 1
    finally:
        torch.randn = real
    got, want = np.asarray(log), g["losses"]
    assert np.allclose(got[:, 0], want[:, 0], rtol=REL) and np.allclose(got[:, 2], want[:, 2], rtol=REL), (got, want)
    assert np.allclose(got[:, 1], want[:, 1], rtol=10 * REL, atol=1e-7)
    sd, g0 = model.state_dict(), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        assert_update_close(k, sd[k].cpu().numpy(), v, init[k], s, 1e-4, True, g0.get(k))
    torch.randn = lambda *a, **k: torch.from_numpy(g["eps_test"][:, :zc].copy()).to(k.get("device", "cpu"))
    try:
        preds = model.pred_recon([pd.DataFrame(g[f"x{i}"]) for i in range(len(dims))], None, None, "poe")
    finally:
        torch.randn = real
    for i in range(len(dims)):
        assert relerr(preds[i], g[f"pred{i}"]) < REL, i
    devs = model.reconstruction_deviation_multimodal([g[f"x{i}"] for i in range(len(dims))], preds)
    assert len(devs) == len(dims) and devs[0].shape == (n,)
    model.close()


def test_train_test_programs_with_model_flag(tmp_path):
    """``-Model`` of the train script (:141-148) on a synthetic HCPimage dataset: mmJSD == cVAE_multimodal with PoE bit for
    bit (the same kernels, `combine` ignored); DMVAE / mmVAEPlus / WeightedDMVAE train, pickle as the drop-in classes
    and go through the test program + group analysis, and so does mvtCAE; an unknown name is refused."""
    import argparse
    import pandas as pd
    from multi_modal_normative_modeling_b200 import cli, synthetic, zoo
    synthetic.write_dataset(str(tmp_path), "HCPimage", n=200, seed=4)
    base = dict(dataset_resourse="HCPimage", hz_para_list=[32, 24, 6], combine=None, procedure="SE-MoE", n_splits=2, epochs=3,
                oversample_percentage=1, single_modality=None, base_learning_rate=1e-4, max_learning_rate=5e-3,
                training_class="nm", ensemble_seeds=1, nmmlp=False)
    a = cli.train_main(argparse.Namespace(model="mmJSD", **base), root=tmp_path)
    poe = dict(base, procedure="SE-PoE")
    b = cli.train_main(argparse.Namespace(model="cVAE_multimodal", **poe), root=tmp_path)
    assert np.array_equal(a, b)                         # mmJSD with -P SE-MoE == cVAE_multimodal with PoE
    with pytest.raises(ValueError, match="not recognized"):
        cli.train_main(argparse.Namespace(model="VAE", **base), root=tmp_path)
    from multi_modal_normative_modeling_b200.cVAE import mvtCAE
    mv = dict(base, procedure="SE-gPoE")
    losses = cli.train_main(argparse.Namespace(model="mvtCAE", **mv), root=tmp_path)
    assert np.isfinite(losses).all() and (losses[:, :, 0] < 50).all()        # total = sum_m (kl + 1e-5 ll + 1e-4 tc): O(1)
    assert type(torch.load(tmp_path / "outputs" / "kfold_analysis" / "supervised_cvae" / "000" / "cVAE_model.pkl", weights_only=False)) is mvtCAE
    cli.test_main(argparse.Namespace(model="mvtCAE", **mv), root=tmp_path)
    assert 0.0 <= cli.analysis_main(argparse.Namespace(model="mvtCAE", **mv), root=tmp_path)[0][2][0] <= 1.0
    for model, cls in (("DMVAE", zoo.DMVAE), ("mmVAEPlus", zoo.mmVAEPlus), ("WeightedDMVAE", zoo.WeightedDMVAE)):
        for hz in ([32, 24, 40], [32, 24, 6]):          # latent 40 > c_dim 29: 11 shared dimensions; latent 6: none (the default situation)
            ns = dict(base, hz_para_list=hz)
            losses = cli.train_main(argparse.Namespace(model=model, **ns), root=tmp_path)
            assert losses.shape == (2, 3, 3) and np.isfinite(losses).all()
            assert (losses[:, :, 1] > 0).all() if hz[-1] > 29 else (losses[:, :, 1] == 0).all()        # kl of the shared part
            saved = torch.load(tmp_path / "outputs" / "kfold_analysis" / "supervised_cvae" / "001" / "cVAE_model.pkl",
                               weights_only=False)
            assert type(saved) is cls and saved.latent_dim == hz[-1]
            cli.test_main(argparse.Namespace(model=model, **ns), root=tmp_path)
            d = tmp_path / "deviation" / "supervised_cvae" / "HCPimage" / "SE-MoE" / "path_model" / "fMRI"
            rec = pd.read_csv(d / "reconstruction_fMRI.csv").iloc[:, 4:].to_numpy()
            nrm = pd.read_csv(d / "normalized_fMRI.csv").iloc[:, 4:].to_numpy()
            roi = pd.read_csv(d / "reconstruction_error_roi_fMRI.csv").iloc[:, 4:].to_numpy()
            assert rec.shape == (200, 116) and rec.min() >= 0.0 and rec.max() <= 1.0          # sigmoid decoders
            assert np.abs((nrm - rec) ** 2 - roi).max() < 1e-4 * max(1.0, np.abs(roi).max())
            summary = cli.analysis_main(argparse.Namespace(model=model, **ns), root=tmp_path)
            assert 0.0 <= summary[0][2][0] <= 1.0
