"""GPU prologue vs the host pandas / sklearn path it replaces: RobustScaler statistics, rank -> qcut
bins, bootstrap / merge gather and row packing -- bit for bit."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prologue_kernels_vs_sklearn_and_pandas():
    from sklearn.preprocessing import RobustScaler
    from multi_modal_normative_modeling_b200 import prologue
    rng = np.random.RandomState(0)
    dev = torch.device("cuda", 0)
    for n_all, n, d in ((1000, 800, 116), (300, 37, 7), (5000, 4097, 20), (64, 1, 3), (9000, 8192, 5)):
        x = rng.randn(n_all, d) * (10.0 ** rng.uniform(0, 3, d)) + rng.randn(d) * 100
        x[:, 0] = np.round(x[:, 0], -1)                                  # ties
        idx = rng.randint(0, n_all, n)                                   # bootstrap: duplicates
        xd, idd = torch.from_numpy(x).to(dev), torch.from_numpy(idx.astype(np.int32)).to(dev)
        center, scale = prologue.robust_fit(xd, idd)
        sc = RobustScaler().fit(x[idx])
        assert np.array_equal(center.cpu().numpy(), sc.center_) and np.array_equal(scale.cpu().numpy(), sc.scale_), (n, d)
        age = rng.randint(22, 37, n_all).astype(np.float64)
        for q in (27, 2):
            if n < q:
                continue
            want = np.asarray(pd.qcut(pd.Series(age[idx]).rank(method="first"), q=q, labels=list(range(q))).values, dtype=np.int64)
            got = prologue.rank_bins(torch.from_numpy(age).to(dev), idd, q).cpu().numpy()
            assert np.array_equal(got, want), (n, q)
    with pytest.raises(RuntimeError, match="8192"):
        prologue.robust_fit(torch.zeros(9000, 2, dtype=torch.float64, device=dev),
                            torch.arange(8193, dtype=torch.int32, device=dev))


def test_gpu_folds_equal_host_folds_bit_for_bit():
    """prepare_folds_gpu == pipeline.prepare_folds + pack_rows for every fold and modality (feature files in a different
    row order than the demographics, bootstrap duplicates, early-fusion table)."""
    from multi_modal_normative_modeling_b200 import pack_rows, pipeline, prologue, synthetic, workloads
    dev = torch.device("cuda", 0)
    data = synthetic.make_hcpimage(400, 24, seed=7)
    subjects = data["subjects"]
    feats, cols = {}, {}
    for k, (name, x) in enumerate(data["features"].items()):
        df = pd.DataFrame(x, columns=[f"r{i}" for i in range(24)])
        df.insert(0, "IID", subjects["IID"].to_numpy())
        feats[name] = df.sample(frac=1.0, random_state=k).reset_index(drop=True)
        cols[name] = [f"r{i}" for i in range(24)]
    host = pipeline.prepare_folds(subjects, feats, cols, hc_label=1, n_splits=4)
    gpu = prologue.prepare_folds_gpu(subjects, feats, cols, 1, dev, n_splits=4)
    assert len(host) == len(gpu) == 4
    for fd, gf in zip(host, gpu):
        for name in feats:
            # the host keeps ONE covariate table per fold (the last modality's, like the reference); with differently
            # ordered feature files the per-modality tables differ, so the host rows are packed with their own covariates
            sc_tr = torch.from_numpy(fd.train_x[name]).to(dev)
            assert torch.equal(gf.train[name][:, :24], sc_tr), name                      # scaled features, bit for bit
            assert torch.equal(gf.test[name][:, :24], torch.from_numpy(fd.test_x[name]).to(dev)), name
        last = list(feats)[-1]
        want_tr = pack_rows(torch.from_numpy(fd.train_x[last]).to(dev), torch.from_numpy(fd.train_c).to(dev))
        want_te = pack_rows(torch.from_numpy(fd.test_x[last]).to(dev), torch.from_numpy(fd.test_c).to(dev))
        assert torch.equal(gf.train[last], want_tr) and torch.equal(gf.test[last], want_te)
    # the ensemble built on the GPU prologue trains to the same bits as the one built on the host pipeline
    hw = workloads.build_host_workload(n_subjects=300, n_splits=3)
    from multi_modal_normative_modeling_b200 import EnsembleTrainer
    a = workloads.to_device(hw, dev, n_seeds=1)
    b = workloads.to_device(hw, dev, n_seeds=1, gpu_prologue=True)
    for k in a.packed:
        assert torch.equal(a.packed[k], b.packed[k]), k
    ta, tb = EnsembleTrainer(a.specs, device=dev), EnsembleTrainer(b.specs, device=dev)
    ta.train_steps(4); tb.train_steps(4)
    torch.cuda.synchronize()
    assert torch.equal(ta.params, tb.params)
    ta.close(); tb.close()
