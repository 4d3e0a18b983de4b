"""GPU tests of the drop-in boundary beyond the training loop: encode / decode / reparameterise / combine_latent of
both classes against calls recorded from the reference (tests/golden/pieces_M3.npz), the per-step API over a RAGGED
epoch (Adam moments must survive the change of minibatch size), the nmmlp program end to end, the latent-space
normative deviation kernel, classification metrics against the sklearn recordings, whole-epoch training of members
with different fold sizes."""
import argparse
import os

import numpy as np
import pandas as pd
import pytest
import torch

from helpers import assert_losses_close, assert_update_close, load, loop_batches, relerr, sub

pytestmark = pytest.mark.gpu
REL = 1e-4


def test_pieces_encode_decode_combine_vs_reference(golden_dir):
    from multi_modal_normative_modeling_b200.cVAE import cVAE, cVAE_multimodal
    g = load(golden_dir, "pieces_M3")
    dims = [int(d) for d in g["dims"]]
    model = cVAE_multimodal(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), modalities=3, non_linear=True)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "init/").items()})
    model.to("cuda")
    c = torch.from_numpy(g["c"]).cuda()
    z = torch.from_numpy(g["zin"]).cuda()
    mus, lvs = [], []
    for m in range(3):
        mu, lv = model.encode(torch.from_numpy(g[f"x{m}"]).cuda(), c, m)
        assert relerr(mu.cpu().numpy(), g[f"enc_mu{m}"]) < REL and relerr(lv.cpu().numpy(), g[f"enc_lv{m}"]) < REL
        d = model.decode(z, c, m)
        assert relerr(d.loc.cpu().numpy(), g[f"dec{m}"]) < REL
        assert relerr(d.scale.cpu().numpy(), g[f"dec_scale{m}"]) < 1e-6
        mus.append(mu); lvs.append(lv)
    mus, var = torch.stack(mus), torch.exp(torch.stack(lvs))
    for comb in ("PoE", "gPoE", "MoE", "MoPoE", "mopoe", "GPOE"):
        key = {"mopoe": "MoPoE", "GPOE": "gPoE"}.get(comb, comb)
        mu_c, var_c = model.combine_latent(mus, var, comb)
        assert relerr(mu_c.detach().cpu().numpy(), g[f"comb_mu/{key}"]) < REL
        assert relerr(var_c.detach().cpu().numpy(), g[f"comb_var/{key}"]) < REL
    with pytest.raises(ValueError, match="No such combination method"):
        model.combine_latent(mus, var, "concat")
    torch.manual_seed(5)
    e = torch.randn_like(mus[0])
    torch.manual_seed(5)
    assert torch.allclose(model.reparameterise(mus[0], lvs[0]), mus[0] + e * torch.exp(0.5 * lvs[0]))
    single = cVAE(13, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), non_linear=True)
    single.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "sinit/").items()})
    single.to("cuda")
    mu, lv = single.encode(torch.from_numpy(g["x0"]).cuda(), c)
    assert relerr(mu.cpu().numpy(), g["s_enc_mu"]) < REL and relerr(lv.cpu().numpy(), g["s_enc_lv"]) < REL
    assert relerr(single.decode(z, c).loc.cpu().numpy(), g["s_dec"]) < REL
    model.close(); single.close()


@pytest.mark.parametrize("name", ["nmmlp_M2_small", "mm_M1_L4"])
def test_module_per_step_api_over_a_ragged_epoch(golden_dir, name):
    """The reference's loop body on the drop-in module over full AND partial batches: the Adam moments live in the
    module (not in a per-batch-size engine), so the trajectory matches the reference's recording."""
    from multi_modal_normative_modeling_b200.cVAE import cVAE_multimodal, cVAE_multimodal_endtoend
    g = load(golden_dir, name)
    nmmlp = name.startswith("nmmlp")
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    cls = cVAE_multimodal_endtoend if nmmlp else cVAE_multimodal
    model = cls(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), learning_rate=1e-4,
                modalities=len(dims), non_linear=True).to("cuda")
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    c = torch.from_numpy(g["c"]).cuda()
    n, b, epochs = int(g["n"]), int(g["batch"]), int(g["epochs"])
    real_randn = torch.randn
    losses, s = [], 0
    for _ in range(epochs):
        for r0, rows in loop_batches(n, b):
            torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows]).to(k.get("device", "cpu"))
            try:
                fwd = model.forward_multimodal([x[r0:r0 + rows] for x in xs], [c[r0:r0 + rows]] * len(dims), str(g["combine"]))
            finally:
                torch.randn = real_randn
            loss = model.loss_function_multimodal(xs, fwd, None) if nmmlp else model.loss_function_multimodal(xs, fwd)
            model.optimizer1.zero_grad()
            loss["total"].backward()
            model.optimizer1.step()
            losses.append([float(loss["total"]), float(loss["kl"]), float(loss["ll"])])
            s += 1
    assert_losses_close(np.array(losses), g["losses"], REL)
    init, g0 = sub(g, "init/"), sub(g, "grad/")
    for k, v in sub(g, "final/").items():   # the UPDATE: a reset of the Adam moments would show up as a 3x-lr spike
        assert_update_close(k, model.state_dict()[k].cpu().numpy(), v, init[k], len(losses), 1e-4, False, g0.get(k))
    st = model.optimizer1.state_dict()["state"]
    assert len(st) > 0 and int(next(iter(st.values()))["step"]) == len(losses)
    assert len(model._cache()) == 2                      # one engine per minibatch size, none rebuilt
    with pytest.raises(ValueError):
        model.loss_function_multimodal(xs, {"mu_multimodal": fwd["mu_multimodal"].clone()})
    model.close()


def test_train_epochs_with_unequal_fold_sizes_and_lr_schedule_bounds():
    """Members with DIFFERENT steps per epoch in one launch: each takes exactly epochs x its own steps (ADVICE r1:
    over-training + out-of-bounds lr_steps read); a schedule shorter than the requested steps is refused."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    from oracle import cvae_torch
    rng = np.random.RandomState(4)
    d, c_dim, z, batch, epochs = 12, 5, 3, 8, 3
    specs, refs = [], []
    for k, n in enumerate((23, 8, 17, 40)):
        x = rng.randn(n, d).astype(np.float32)
        c = np.zeros((n, c_dim), np.float32); c[np.arange(n), rng.randint(0, c_dim, n)] = 1
        spe = -(-n // batch)
        lr = (1e-4 * (1 + 0.3 * np.cos(np.arange(epochs * spe)))).astype(np.float32)
        torch.manual_seed(20 + k)
        model = cvae_torch.OracleCVAEMultimodal([d], [10, 7], z, c_dim, 1e-4, 1, True)
        sd = {a: b.detach().clone() for a, b in model.state_dict().items()}
        specs.append(MemberSpec([d], [10, 7], z, c_dim, [pack_rows(torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda())],
                                batch=batch, seed=100 + k, state_dict=sd, lr_steps=torch.from_numpy(lr).cuda()))
        refs.append((model, x, c, lr, spe))
    for flags in (0, 8, 16):                      # pipelined, FP32, generic engines
        tr = EnsembleTrainer(specs)
        losses = tr.train_epochs(epochs, record_losses=True, flags=flags).cpu().numpy()
        done = tr.steps_done()
        for k, (model, x, c, lr, spe) in enumerate(refs):
            assert int(done[k]) == epochs * spe, (k, done)
            assert np.isfinite(losses[k, :epochs * spe]).all() and np.isnan(losses[k, epochs * spe:]).all()
        # the in-kernel eps stream is seeded per member: replay it through the oracle loop
        from multi_modal_normative_modeling_b200 import scoring
        k = 0
        model, x, c, lr, spe = refs[k]
        torch.manual_seed(20 + k)
        model = cvae_torch.OracleCVAEMultimodal([d], [10, 7], z, c_dim, 1e-4, 1, True)
        log = cvae_torch.reference_train_loop(
            model, [torch.from_numpy(x)], [torch.from_numpy(c).long()], "poe", epochs, batch,
            eps_fn=lambda s, rows: scoring.philox_normal(100 + k, s, batch * z, 0).view(batch, z)[:rows].cpu(),
            lr_fn=lambda step: float(lr[step - 1]))
        assert np.allclose(losses[k, :epochs * spe], log, rtol=2e-4, atol=1e-5)
        with pytest.raises(RuntimeError, match="lr_steps"):
            tr.train_epochs(1)                    # the schedules are exhausted
        with pytest.raises(RuntimeError, match="lr_steps"):
            tr.train_steps(1)
        tr.close()


def test_latent_deviation_kernel_vs_reference(golden_dir):
    from multi_modal_normative_modeling_b200 import scoring
    g = load(golden_dir, "latent_deviation")
    tags = "abcd"
    z, dev = scoring.latent_deviation([torch.from_numpy(g[f"{t}/mu_train"]).cuda() for t in tags],
                                      [torch.from_numpy(g[f"{t}/mu"]).cuda() for t in tags],
                                      [torch.from_numpy(g[f"{t}/logvar"]).cuda() for t in tags])
    torch.cuda.synchronize()
    for i, t in enumerate(tags):
        assert relerr(z[i].cpu().numpy(), g[f"{t}/sep"]) < 1e-5, t
        assert relerr(dev[i].cpu().numpy(), g[f"{t}/dev"]) < 1e-5, t


def test_classification_performance_vs_sklearn_recordings(golden_dir):
    """cli.classification_performance (GPU pair-count AUC + Youden threshold) vs the values recorded from sklearn at
    the reference's call sites (group analysis :105-157)."""
    from multi_modal_normative_modeling_b200 import cli
    g = load(golden_dir, "host_callsites")
    for tag in ("a", "b"):
        auc, acc, sens, spec, ratio = cli.classification_performance(g[f"roc/{tag}/scores"], g[f"roc/{tag}/labels"],
                                                                      torch.device("cuda", 0))
        assert abs(auc - float(g[f"roc/{tag}/auc"])) < 1e-12
        assert acc == float(g[f"roc/{tag}/acc"]) and sens == float(g[f"roc/{tag}/sens"]) and spec == float(g[f"roc/{tag}/spec"])
        assert abs(ratio - auc / (1 - auc)) < 1e-12


def test_nmmlp_program_end_to_end(tmp_path):
    """multimodal_kfold_cvae_nmmlp.py all: HC-only training rows, -MSE term, cyclic LR, diagnosis files, metrics."""
    from multi_modal_normative_modeling_b200 import cli, synthetic
    synthetic.write_dataset(str(tmp_path), "HCPimage", n=300, seed=5)
    out = cli.nmmlp_main(["all", "-R", "HCPimage", "-H", "32", "24", "6", "-P", "SE-gPoE", "-E", "4", "-K", "3"], root=tmp_path)
    assert out["losses"].shape == (3, 4, 3) and np.isfinite(out["losses"]).all()
    assert (out["losses"][:, :, 2] <= 0).all()                      # ll = -MSE
    md = tmp_path / "outputs" / "kfold_analysis" / "supervised_cvae"
    m = torch.load(md / "000" / "cVAE_model.pkl", weights_only=False)
    assert type(m).__name__ == "cVAE_multimodal_endtoend" and any(k.startswith("mlp.") for k in m.state_dict())
    diag = pd.read_csv(md / "001" / "diagnosis_results.csv")
    assert list(diag.columns) == ["participant_id", "Diagnosis", "True_Label"] and set(diag["True_Label"]) <= {0, 1}
    errs = [pd.read_csv(md / "001" / n / f"reconstruction_error_{n}.csv")["Reconstruction error"].to_numpy()
            for n in ("T1w_sMRI", "T2w_sMRI", "fMRI")]
    assert np.allclose(diag["Diagnosis"].to_numpy(), np.mean(errs, axis=0), rtol=1e-5)
    assert out["metrics"].shape == (3, 5) and (out["metrics"][:, 0] >= 0).all() and (out["metrics"][:, 0] <= 1).all()
    assert (tmp_path / "outputs" / "analysis_results" / "performance_metrics.txt").exists()
    # the training rows were healthy controls only: the schedule length = epochs * ceil(n_hc_rows / 256)
    tr_ids = pd.read_csv(tmp_path / "outputs" / "kfold_analysis" / "train_ids_000.csv")
    assert len(tr_ids) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["default", "fp32"])
def test_mmjsd_baseline_dropin_vs_reference(golden_dir, engine):
    """f4: drop-in ``mmJSD`` through the reference loop (fwd -> loss -> zero_grad -> backward -> optimizer1.step()) with
    `combine="moe"` passed in -- and ignored, like the reference -- against the recording of the unmodified class."""
    import cVAE as shim
    from helpers import assert_update_close, load, relerr, sub
    g = load(golden_dir, "mmjsd_M3")
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    model = shim.mmJSD(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), learning_rate=1e-4, modalities=len(dims),
                       non_linear=True)
    init = sub(g, "init/")
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), init[k]), k
    model.to("cuda")
    if engine == "fp32":
        from multi_modal_normative_modeling_b200 import _lib
        model._engine_flags = _lib.TRAIN_FP32
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    c = torch.from_numpy(g["c"]).long().cuda()
    n, b = int(g["n"]), int(g["batch"])
    real = torch.randn
    log, s = [], 0
    try:
        for _ in range(int(g["epochs"])):
            for r0 in range(0, n, b):
                rows = min(b, n - r0)
                torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s][:rows]).to(k.get("device", "cpu"))
                fwd = model.forward_multimodal([x[r0:r0 + rows] for x in xs], [c[r0:r0 + rows]] * len(dims), "moe")
                torch.randn = real
                loss = model.loss_function_multimodal([x[r0:r0 + rows] for x in xs], fwd)
                if s == 0:
                    assert relerr(fwd["mu_multimodal"].detach().cpu().numpy(), g["mu"]) < 1e-4
                    assert float(model.multimodal_jsd([fwd["mu_multimodal"].detach()] * 3, [fwd["logvar_multimodal"].detach()] * 3)) == 0.0
                model.optimizer1.zero_grad()
                loss["total"].backward()
                model.optimizer1.step()
                log.append([float(loss[k].detach()) for k in ("total", "kl", "ll")])
                s += 1
    finally:
        torch.randn = real
    got, want = np.asarray(log), g["losses"]
    assert np.allclose(got[:, 0], want[:, 0], rtol=1e-4) and np.allclose(got[:, 2], want[:, 2], rtol=1e-4)
    assert np.allclose(got[:, 1], want[:, 1], rtol=1e-3)
    sd, g0 = model.state_dict(), sub(g, "grad/")
    for k, v in sub(g, "final/").items():
        # FP32 engine: the reference's trajectory to 2e-4 of the largest update on 99 % of the elements (measured 1e-5).
        # Default engine (BF16x3; batch 128 -> the generic one): after the first Adam step one hidden unit of encoder 0 sits
        # within rounding distance of the leaky-relu kink for one of the 22 samples of the ragged batch and takes the other
        # branch (tests/tools/diag_traj.py: that step's gradient agrees with the FP32 engine to 4e-5 in the median and differs by
        # 6e-2 in ONE row of encoder 0 / layer 1, hence by a rank-one term in layer 0).  The fixture's knife-edge scan
        # (oracle/make_golden.py::clean_seed) covers the initial weights only, so the looser bound applies here.
        assert_update_close(k, sd[k].cpu().numpy(), v, init[k], s, 1e-4, engine == "fp32", g0.get(k), q99_tc=6e-2, mean_tc=1e-2)
    import pandas as pd
    torch.randn = lambda *a, **k: torch.from_numpy(g["eps_test"])
    try:
        preds = model.pred_recon([pd.DataFrame(g[f"x{i}"].astype(np.float64)) for i in range(len(dims))], g["c"], None, "moe")
    finally:
        torch.randn = real
    for i in range(len(dims)):
        assert relerr(preds[i], g[f"pred{i}"]) < 2e-4, i
    model.close()
