"""End-to-end GPU tests: k-fold training + deviation scoring against the CPU oracle on synthetic
HCP-shaped data (|dAUC| <= 0.01, north_star), the multimodal_kfold_* CLI file contract, and the
drop-in nn.Module per-step API against the reference's golden vectors."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_kfold_train_and_score_vs_oracle_auc():
    """cfg1 shape (single modality, 5 folds): identical weights and identical eps draws in the oracle
    (eager PyTorch restatement of the reference loop) and in the fused kernels; final per-subject
    deviation AUC, per-ROI AUC and HC-referenced z-scores must agree (|dAUC| <= 0.01)."""
    from oracle import cvae_torch, deviation as odev
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, scoring, workloads
    hw = workloads.build_host_workload(n_subjects=400, d=150, n_splits=5, early_fusion=False, hidden=(110, 110))
    name, epochs, batch = "T1w_sMRI", 6, 256
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(0)
    specs, oracle_out, eps_all, eps_test_all = [], [], [], []
    for fd in hw.folds:
        x, c = fd.train_x[name], fd.train_c
        spe = -(-x.shape[0] // batch)
        torch.manual_seed(42)
        model = cvae_torch.OracleCVAEMultimodal([150], [110, 110], 10, 29, 1e-4, 1, True)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        eps = rng.randn(epochs * spe, batch, 10).astype(np.float32)
        eps_test = rng.randn(fd.test_x[name].shape[0], 10).astype(np.float32)
        cvae_torch.reference_train_loop(model, [torch.from_numpy(x)], [torch.from_numpy(c).long()], "gPoE", epochs, batch,
                                        eps_fn=lambda s, rows, e=eps: torch.from_numpy(e[s][:rows]))
        xt, ct = torch.from_numpy(fd.test_x[name]), torch.from_numpy(fd.test_c).long()
        pred = model.pred_recon([xt], ct, "gPoE", torch.from_numpy(eps_test))[0].numpy()
        pred_tr = model.pred_recon([torch.from_numpy(x)], torch.from_numpy(c).long(), "gPoE",
                                   torch.zeros(x.shape[0], 10))[0].numpy()
        hc = fd.train_df["DIA"].to_numpy() == 1
        mean, std = odev.normative_stats(odev.recon_deviation_roi(x, pred_tr)[hc])
        roi = odev.recon_deviation_roi(fd.test_x64[name], pred)
        labels = (fd.test_df["DIA"].to_numpy() != 1).astype(np.uint8)
        oracle_out.append(dict(subj=odev.recon_deviation(fd.test_x64[name], pred), z=odev.zscores(roi, mean, std),
                               labels=labels, hc=hc))
        xc = [pack_rows(torch.from_numpy(x).to(dev), torch.from_numpy(c).to(dev))]
        specs.append(MemberSpec([150], [110, 110], 10, 29, xc, batch=batch, state_dict=sd))
        eps_all.append(eps); eps_test_all.append(eps_test)
    tr = EnsembleTrainer(specs, device=dev)
    tr.train_steps(eps_all[0].shape[0], eps=torch.from_numpy(np.stack(eps_all)).to(dev))
    test_xc = [[pack_rows(torch.from_numpy(fd.test_x[name]).to(dev), torch.from_numpy(fd.test_c).to(dev))]
               for fd in hw.folds]
    xhat, _, _ = tr.reconstruct(test_xc, mode="sample", eps=[torch.from_numpy(e).to(dev) for e in eps_test_all])
    xhat_tr, _, _ = tr.reconstruct([s.xc for s in specs], mode="mean")
    stats = scoring.normative_stats([s.xc[0] for s in specs], [h[0] for h in xhat_tr],
                                    [torch.from_numpy(o["hc"].astype(np.uint8)).to(dev) for o in oracle_out])
    roi, z, subj = scoring.deviation([t[0] for t in test_xc], [h[0] for h in xhat], stats)
    labels = [torch.from_numpy(o["labels"]).to(dev) for o in oracle_out]
    subj_auc = scoring.auc(subj, labels)
    roi_auc = scoring.auc(z, labels)
    torch.cuda.synchronize()
    for f, o in enumerate(oracle_out):
        want = odev.auc(o["subj"], o["labels"])
        assert abs(float(subj_auc[f][0]) - want) <= 0.01, (f, float(subj_auc[f][0]), want)
        assert relerr(subj[f].cpu().numpy(), o["subj"]) < 1e-3
        zz = z[f].cpu().numpy()
        assert np.abs(zz - o["z"]).max() < 2e-2 * max(1.0, np.abs(o["z"]).max())       # z tolerance: 2 % of range
        for col in range(0, 150, 13):
            assert abs(float(roi_auc[f][col]) - odev.auc(o["z"][:, col], o["labels"])) <= 0.01
    tr.close()


def test_cli_file_contract(tmp_path):
    """train -> test -> group analysis on a synthetic HCPimage dataset, reference file layout."""
    import argparse
    from multi_modal_normative_modeling_b200 import cli, synthetic
    synthetic.write_dataset(str(tmp_path), "HCPimage", n=240, seed=3)
    ns = dict(dataset_resourse="HCPimage", hz_para_list=[32, 24, 6], combine=None, procedure="SE-gPoE", n_splits=3,
              epochs=3, oversample_percentage=1, model="cVAE_multimodal", single_modality=None,
              base_learning_rate=1e-4, max_learning_rate=5e-3, training_class="nm", ensemble_seeds=2, nmmlp=False)
    losses = cli.train_main(argparse.Namespace(**ns), root=tmp_path)
    assert losses.shape == (6, 3, 3) and np.isfinite(losses).all()
    kf = tmp_path / "outputs" / "kfold_analysis"
    assert (kf / "train_ids_002.csv").exists() and (kf / "supervised_cvae" / "001" / "cVAE_model.pkl").exists()
    cli.test_main(argparse.Namespace(**ns), root=tmp_path)
    for name in ("T1w_sMRI", "T2w_sMRI", "fMRI"):
        d = tmp_path / "deviation" / "supervised_cvae" / "HCPimage" / "SE-gPoE" / "path_model" / name
        nrm = pd.read_csv(d / f"normalized_{name}.csv")
        rec = pd.read_csv(d / f"reconstruction_{name}.csv")
        roi = pd.read_csv(d / f"reconstruction_error_roi_{name}.csv")
        err = pd.read_csv(d / f"reconstruction_error_{name}.csv")
        fi = pd.read_csv(d / f"deviation_as_feature_importance_{name}.csv")
        assert len(nrm) == 240 and list(nrm.columns[:4]) == ["participant_id", "DIA", "AGE", "PTGENDER"]
        a, b, r = nrm.iloc[:, 4:].to_numpy(), rec.iloc[:, 4:].to_numpy(), roi.iloc[:, 4:].to_numpy()
        # the stored-CSV identities of the reference's own artefacts
        assert np.abs((a - b) ** 2 - r).max() < 1e-4 * max(1.0, np.abs(r).max())
        assert np.abs(r.mean(1) - err["Reconstruction error"].to_numpy()).max() < 1e-4 * max(1.0, r.mean(1).max())
        assert list(fi.columns[4:]) == [str(i) for i in range(1, 117)] and np.array_equal(fi.iloc[:, 4:].to_numpy(), r)
    summary = cli.analysis_main(argparse.Namespace(**ns), root=tmp_path)
    assert len(summary) == 1 and 0.0 <= summary[0][2][0] <= 1.0
    aucs = np.loadtxt(tmp_path / "cvae_auc_and_std.csv", delimiter=",")
    assert len(aucs) == 4 and abs(np.std(aucs[:-1]) - aucs[-1]) < 1e-15


@pytest.mark.parametrize("name", ["mm_M1_small", "mm_M3_gpoe", "mm_M4_mopoe"])
def test_module_per_step_api_vs_reference(golden_dir, name):
    """The reference's own loop body on the drop-in module: forward_multimodal -> loss ->
    zero_grad -> backward -> optimizer1.step(), with the reference's eps draws injected through
    torch.randn (as the reference itself consumes them)."""
    from multi_modal_normative_modeling_b200.cVAE import cVAE_multimodal
    g = load(golden_dir, name)
    dims = [int(d) for d in g["dims"]]
    torch.manual_seed(int(g["seed"]))
    model = cVAE_multimodal(dims, [int(h) for h in g["hidden"]], int(g["z"]), int(g["c_dim"]), learning_rate=1e-4,
                            modalities=len(dims), non_linear=True)
    for k, v in model.state_dict().items():                        # seed-exact initialisation
        assert np.array_equal(v.numpy(), g["init/" + k]), k
    model.to("cuda")
    xs = [torch.from_numpy(g[f"x{i}"]).cuda() for i in range(len(dims))]
    cs = [torch.from_numpy(g["c"]).long().cuda() for _ in dims]
    real_randn = torch.randn
    losses = []
    for s in range(g["eps"].shape[0]):
        torch.randn = lambda *a, **k: torch.from_numpy(g["eps"][s]).to(k.get("device", "cpu"))
        try:
            fwd = model.forward_multimodal(xs, cs, str(g["combine"]))
        finally:
            torch.randn = real_randn
        loss = model.loss_function_multimodal(xs, fwd)
        model.optimizer1.zero_grad()
        loss["total"].backward()
        if s == 0:
            assert relerr(fwd["mu_multimodal"].detach().cpu().numpy(), g["mu"]) < 1e-4
            assert relerr(fwd["x_recons"][0].loc.detach().cpu().numpy(), g["xrecon0"]) < 1e-4
            for k, p in model.named_parameters():
                if "grad/" + k in g:
                    assert relerr(p.grad.cpu().numpy(), g["grad/" + k]) < 1e-4, k
        model.optimizer1.step()
        losses.append([float(loss["total"]), float(loss["kl"]), float(loss["ll"])])
    assert np.allclose(np.array(losses), g["losses"], rtol=1e-4)
    for k, v in model.state_dict().items():
        assert relerr(v.cpu().numpy(), g["final/" + k]) < 1e-5, k
    with pytest.raises(ValueError, match="No such combination method"):
        model.forward_multimodal(xs, cs, "concat")


def test_deviation_scorer_equals_the_per_call_api():
    """DeviationScorer (buffers + argument tables prepared once, six launches per pass) == the per-call
    scoring API, bit for bit, on a heterogeneous ensemble with ragged / unaligned segment sizes."""
    import numpy as np
    import torch
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, scoring
    rng = np.random.RandomState(3)
    specs, train, test, masks, labels = [], [], [], [], []
    for k, (dims, n_tr, n_te) in enumerate([([17], 41, 13), ([9, 6], 30, 9), ([116], 64, 20), ([7], 25, 11)]):
        c_tr = torch.zeros(n_tr, 5).cuda(); c_tr[:, k % 5] = 1
        c_te = torch.zeros(n_te, 5).cuda(); c_te[:, k % 5] = 1
        tr_x = [pack_rows(torch.from_numpy(rng.randn(n_tr, d).astype(np.float32)).cuda(), c_tr) for d in dims]
        te_x = [pack_rows(torch.from_numpy(rng.randn(n_te, d).astype(np.float32)).cuda(), c_te) for d in dims]
        specs.append(MemberSpec(dims, [12, 10], 4, 5, tr_x, batch=16, seed=k))
        train.append(tr_x); test.append(te_x)
        m = rng.rand(n_tr) < 0.7; m[:2] = True
        masks.append(torch.from_numpy(m.astype(np.uint8)).cuda())
        lab = (rng.rand(n_te) < 0.4).astype(np.uint8); lab[0], lab[-1] = 0, 1
        labels.append(torch.from_numpy(lab).cuda())
    torch.manual_seed(0)
    tr = EnsembleTrainer(specs)
    tr.params.normal_(0, 0.1)
    tr.train_steps(3)
    sc = scoring.DeviationScorer(tr, train, test, masks, labels).run()
    hat_tr, _, _ = tr.reconstruct(train, mode="mean")
    hat_te, _, _ = tr.reconstruct(test, mode="mean")
    s = 0
    for i, sp in enumerate(specs):
        for k in range(len(sp.input_dims)):
            stats = scoring.normative_stats([train[i][k]], [hat_tr[i][k]], [masks[i]])
            roi, z, subj = scoring.deviation([test[i][k]], [hat_te[i][k]], stats)
            assert torch.equal(sc.seg_stats(s), stats[0])
            assert torch.equal(sc.seg_roi(s), roi[0]) and torch.equal(sc.seg_z(s), z[0]) and torch.equal(sc.seg_subj(s), subj[0])
            assert torch.equal(sc.seg_auc_roi(s), scoring.auc(z, [labels[i]])[0])
            assert torch.equal(sc.auc_subj[s:s + 1], scoring.auc(subj, [labels[i]])[0])
            s += 1
    rec = sc.member_records()
    assert rec.shape == (sc.n_seg, 1 + 3 * 116)
    assert torch.equal(rec[3, 1:117], sc.seg_stats(3)[0].double()) and torch.equal(rec[:, 0], sc.auc_subj)
    assert torch.equal(rec[0, 1 + 2 * 116:1 + 2 * 116 + 17], sc.seg_auc_roi(0)) and float(rec[0, 1 + 17:1 + 116].abs().max()) == 0
    # the all-gather payload in one launch (nmb_member_records): the same record + the NaN-padded per-subject deviations
    nmax = max(sc.n_test) + 3
    tbl = sc.member_table(d_max=120, n_test_max=nmax)
    wide = sc.member_records(d_max=120)
    assert tbl.shape == (sc.n_seg, 1 + 3 * 120 + nmax) and torch.equal(tbl[:, :1 + 3 * 120], wide)
    for s_ in range(sc.n_seg):
        n = sc.n_test[s_]
        assert torch.equal(tbl[s_, 1 + 360:1 + 360 + n], sc.seg_subj(s_).double()) and bool(torch.isnan(tbl[s_, 1 + 360 + n:]).all())
    tr.close()


def _kfold_auc_vs_oracle(dims_names, combine, early_fusion, production_rng, d=116, n_subjects=400, epochs=6):
    """5-fold train + deviation scoring of ONE configuration on the GPU and with the CPU oracle (same initial weights,
    same eps draws): per-fold AUC of the modality-averaged per-subject deviation (group analysis :212-215) and per-ROI
    z-score AUCs must agree within |dAUC| <= 0.01 (north_star).  production_rng: the kernels draw eps themselves (Philox
    stream of oracle/philox.py) instead of taking injected draws; the oracle replays that stream."""
    from oracle import cvae_torch, deviation as odev, philox
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, scoring, workloads
    hw = workloads.build_host_workload(n_subjects=n_subjects, d=d, n_splits=5, early_fusion=early_fusion, hidden=(110, 110))
    names = [n for n in hw.names if n in dims_names] if dims_names else [workloads.EARLY]
    dims = [hw.dims[n] for n in names]
    M, z, batch = len(names), 10, 256
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(1)
    specs, oracle_out, eps_all, eps_test_all, test_xc = [], [], [], [], []
    for f, fd in enumerate(hw.folds):
        xs = [fd.train_x[n] for n in names]
        c = fd.train_c
        n_tr, n_te = xs[0].shape[0], fd.test_x[names[0]].shape[0]
        spe = -(-n_tr // batch)
        seed = 5000 + f
        torch.manual_seed(42)
        model = cvae_torch.OracleCVAEMultimodal(dims, [110, 110], z, 29, 1e-4, M, True)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        if production_rng:
            eps = np.stack([philox.normals(seed, s, batch * z, 0).reshape(batch, z) for s in range(epochs * spe)])
            tiles = -(-n_te // 256)
            eps_test = np.concatenate([philox.normals(seed, t, 256 * z, 1).reshape(256, z) for t in range(tiles)])[:n_te]
        else:
            eps = rng.randn(epochs * spe, batch, z).astype(np.float32)
            eps_test = rng.randn(n_te, z).astype(np.float32)
        cvae_torch.reference_train_loop(model, [torch.from_numpy(x) for x in xs], [torch.from_numpy(c).long()] * M, combine,
                                        epochs, batch, eps_fn=lambda s, rows, e=eps: torch.from_numpy(e[s][:rows]))
        ct = torch.from_numpy(fd.test_c).long()
        preds = model.pred_recon([torch.from_numpy(fd.test_x[n]) for n in names], ct, combine, torch.from_numpy(eps_test))
        preds_tr = model.pred_recon([torch.from_numpy(x) for x in xs], torch.from_numpy(c).long(), combine,
                                    torch.zeros(n_tr, z))
        hc = fd.train_df["DIA"].to_numpy() == 1
        labels = (fd.test_df["DIA"].to_numpy() != 1).astype(np.uint8)
        subj, zs = [], []
        for m, n in enumerate(names):
            p, ptr = preds[m].numpy(), preds_tr[m].numpy()
            mean, std = odev.normative_stats(odev.recon_deviation_roi(xs[m], ptr)[hc])
            subj.append(odev.recon_deviation(fd.test_x64[n], p))
            zs.append(odev.zscores(odev.recon_deviation_roi(fd.test_x64[n], p), mean, std))
        oracle_out.append(dict(subj=np.mean(subj, axis=0), z=zs, labels=labels, hc=hc))
        cd = torch.from_numpy(c).to(dev)
        xc = [pack_rows(torch.from_numpy(x).to(dev), cd) for x in xs]
        specs.append(MemberSpec(dims, [110, 110], z, 29, xc, combine=combine, batch=batch, seed=seed, state_dict=sd))
        ctd = torch.from_numpy(fd.test_c).to(dev)
        test_xc.append([pack_rows(torch.from_numpy(fd.test_x[n]).to(dev), ctd) for n in names])
        eps_all.append(eps); eps_test_all.append(eps_test)
    tr = EnsembleTrainer(specs, device=dev)
    assert tr.engine() == "tcgen05-pipelined"
    n_steps = eps_all[0].shape[0]
    if production_rng:
        tr.train_steps(n_steps)
        xhat, _, _ = tr.reconstruct(test_xc, mode="sample")
    else:
        tr.train_steps(n_steps, eps=torch.from_numpy(np.stack(eps_all)).to(dev))
        xhat, _, _ = tr.reconstruct(test_xc, mode="sample", eps=[torch.from_numpy(e).to(dev) for e in eps_test_all])
    xhat_tr, _, _ = tr.reconstruct([s.xc for s in specs], mode="mean")
    for f, o in enumerate(oracle_out):
        mask = torch.from_numpy(o["hc"].astype(np.uint8)).to(dev)
        lab = torch.from_numpy(o["labels"]).to(dev)
        stats = scoring.normative_stats(specs[f].xc, xhat_tr[f], [mask] * M)
        _, zg, subj = scoring.deviation(test_xc[f], xhat[f], stats)
        avg = scoring.mean_rows(subj)                                   # modality averaging on the GPU
        got = float(scoring.auc([avg], [lab])[0][0])
        want = odev.auc(o["subj"], o["labels"])
        assert abs(got - want) <= 0.01, (f, got, want)
        assert relerr(avg.cpu().numpy(), o["subj"]) < 2e-3
        roi_auc = scoring.auc(zg, [lab] * M)
        for m in range(M):
            for col in range(0, dims[m], 17):
                assert abs(float(roi_auc[m][col]) - odev.auc(o["z"][m][:, col], o["labels"])) <= 0.01, (f, m, col)
    tr.close()


def test_cfg2_early_fusion_kfold_auc_vs_oracle():
    """BASELINE configs[1]: early_fusion_modalities (concatenated T1w + T2w + fMRI, D = 348), 5 folds."""
    _kfold_auc_vs_oracle(None, "gPoE", early_fusion=True, production_rng=False)


def test_cfg3_multimodal_fusion_kfold_auc_vs_oracle():
    """BASELINE configs[2]: per-modality encoders / decoders with latent fusion (gPoE, three modalities) and the group
    analysis' modality-averaged deviation AUC."""
    _kfold_auc_vs_oracle(("T1w_sMRI", "T2w_sMRI", "fMRI"), "gPoE", early_fusion=False, production_rng=False)


def test_cfg1_production_philox_draws_kfold_auc_vs_oracle():
    """The production path end to end: in-kernel Philox eps for training and for the sampled test-time z, against the
    oracle replaying the documented stream (oracle/philox.py)."""
    _kfold_auc_vs_oracle(("T1w_sMRI",), "poe", early_fusion=False, production_rng=True, d=150)


def test_cfg5_scale_stress_properties():
    """BASELINE configs[4] at FULL size (100 000 subjects x 1 000 features, 5 folds -> 80 000 training rows, 313
    minibatches per epoch): too large for the CPU oracle, so parity is checked through size-independent properties --
    the pipelined kernel against the FP32 engine on the same rows and draws (per-step losses), bit-determinism of a
    whole epoch, the pipelined forward-only reconstruction against the generic engine on all 20 000 test rows, and
    the multi-tile AUC path (20 000 rows > one 1 024-row tile) against exact pair counting on the host."""
    from oracle import deviation as odev
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, _lib, pack_rows, scoring
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cuda").manual_seed(5)
    n_tr, n_te, d, c_dim, z = 80000, 20000, 1000, 29, 10

    def rows(n):
        x = torch.randn(n, d, device=dev, generator=g)
        c = torch.zeros(n, c_dim, device=dev)
        c[torch.arange(n, device=dev), torch.randint(0, 27, (n,), device=dev, generator=g)] = 1
        c[torch.arange(n, device=dev), 27 + torch.randint(0, 2, (n,), device=dev, generator=g)] = 1
        return pack_rows(x, c)
    train, test = rows(n_tr), rows(n_te)
    from multi_modal_normative_modeling_b200 import workloads
    sd = workloads.init_state_dict(d, (110, 110), z, c_dim, 42)
    mk = lambda: EnsembleTrainer([MemberSpec([d], [110, 110], z, c_dim, [train], seed=9, state_dict=sd),
                                  MemberSpec([d], [110, 110], z, c_dim, [train], seed=10, state_dict=sd)], device=dev)
    a, b, f = mk(), mk(), mk()
    assert a.engine() == "tcgen05-pipelined" and a.steps_per_epoch == [313, 313]
    la = a.train_steps(313, record_losses=True)                   # one full epoch incl. the ragged 128-row last batch
    lb = b.train_steps(313, record_losses=True)
    lf = f.train_steps(12, record_losses=True, flags=_lib.TRAIN_FP32)
    torch.cuda.synchronize()
    assert torch.equal(la, lb) and torch.equal(a.params, b.params)            # bit-deterministic
    assert torch.isfinite(la).all() and float(la[:, -20:, 0].mean()) < float(la[:, :20, 0].mean())   # it trains
    assert np.allclose(la[:, :12].cpu().numpy(), lf.cpu().numpy(), rtol=1e-4)   # same trajectory as the FP32 engine
    xa, mua, _ = a.reconstruct([[test], [test]], mode="mean", want_latent=True)
    xg, mug, _ = a.reconstruct([[test], [test]], mode="mean", want_latent=True, engine="tcs")
    torch.cuda.synchronize()
    assert float((xa[0][0] - xg[0][0]).abs().max()) < 2e-5 * max(1.0, float(xg[0][0].abs().max()))
    assert float((mua[1] - mug[1]).abs().max()) < 2e-5 * max(1.0, float(mug[1].abs().max()))
    _, _, subj = scoring.deviation([test], [xa[0][0]], want_roi=False)
    lab = (torch.rand(n_te, device=dev, generator=g) < 0.3).to(torch.uint8)
    auc, u2 = scoring.auc(subj, [lab], want_pairs=True)
    want, n1, n0 = odev.auc_pairs(subj[0].cpu().numpy(), lab.cpu().numpy())
    assert int(u2[0][0]) == want and float(auc[0][0]) == want / (2.0 * n1 * n0)
    a.close(); b.close(); f.close()
