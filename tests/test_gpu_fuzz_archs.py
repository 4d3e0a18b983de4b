"""Architecture / shape fuzz of the pipelined tcgen05 kernel against the FP32 engine: layer counts and widths,
latent sizes on both sides of the in-register latent path (Z <= 16), reconstruction widths on both sides of the
D <= 128 path, no covariates, several modalities with every fusion, batches that leave the second half ragged
or empty, datasets that end in a partial minibatch, per-step learning rates, both losses."""
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# dims, hidden, latent, c_dim, combine, loss, n_rows, batch
CASES = [
    ([128], [110, 110], 10, 29, "poe", "gauss_ll", 300, 256),
    ([129], [64], 16, 3, "poe", "gauss_ll", 257, 256),
    ([40], [127, 16, 8], 17, 5, "poe", "gauss_ll", 200, 129),
    ([64, 64], [32, 24], 20, 0, "moe", "neg_mse", 150, 128),
    ([7], [8], 1, 1, "poe", "gauss_ll", 19, 8),
    ([200, 30, 5], [50, 40], 6, 4, "gPoE", "gauss_ll", 333, 200),
    ([1000], [110, 110], 10, 29, "poe", "gauss_ll", 600, 256),
    ([116], [110, 110], 32, 29, "poe", "neg_mse", 130, 130),
    ([33, 34, 35, 36], [20], 4, 2, "mopoe", "gauss_ll", 90, 64),
    ([260], [100, 90, 80, 70], 12, 7, "poe", "gauss_ll", 513, 256),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "D%s_H%s_Z%d_C%d_%s_%s_n%d_b%d" % (
    "x".join(map(str, c[0])), "x".join(map(str, c[1])), c[2], c[3], c[4], c[5][:3], c[6], c[7]))
def test_pipelined_engine_tracks_fp32_engine(case):
    from oracle import cvae_torch
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
    dims, hidden, z, c_dim, combine, loss, n, batch = case
    rng = np.random.RandomState(zlib.crc32(str(case).encode()) % (2 ** 31))      # stable across processes
    c = np.zeros((n, c_dim), np.float32)
    if c_dim:
        c[np.arange(n), rng.randint(0, c_dim, n)] = 1
    cc = torch.from_numpy(c).cuda()
    xc = [pack_rows(torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda(), cc) for d in dims]
    spe = -(-n // batch)
    steps = 2 * spe + 1
    lr = torch.from_numpy((1e-4 * (1 + 0.3 * np.cos(np.arange(steps)))).astype(np.float32)).cuda()
    specs = []
    for k in range(3):
        torch.manual_seed(11 + k)
        model = cvae_torch.OracleCVAEMultimodal(dims, hidden, z, c_dim, 1e-4, len(dims), True, loss)
        sd = {a: b.detach().clone() for a, b in model.state_dict().items()}
        specs.append(MemberSpec(dims, hidden, z, c_dim, xc, combine=combine, loss_kind=loss, batch=batch, seed=5 + k,
                                state_dict=sd, lr_steps=lr if k == 1 else None))
    out = {}
    for name, flags in (("fp32", _lib.TRAIN_FP32), ("tc", 0)):
        tr = EnsembleTrainer(specs)
        if name == "tc":
            assert tr.engine() == "tcgen05-pipelined"
        a = tr.train_steps(spe, record_losses=True, flags=flags)
        b = tr.train_steps(steps - spe, record_losses=True, flags=flags)
        torch.cuda.synchronize()
        out[name] = (torch.cat([a, b], 1).cpu().numpy(), tr.params.cpu().numpy().copy())
        tr.close()
    l0, p0 = out["fp32"]
    l1, p1 = out["tc"]
    assert np.isfinite(l1).all() and np.isfinite(p1).all()
    # step 0 runs on identical weights: the per-step bar; later steps also carry Adam's sign-like first updates of
    # near-zero gradients (largest on the 1-row minibatches some cases end with)
    assert np.allclose(l1[:, 0], l0[:, 0], rtol=1e-4, atol=1e-5), float(np.abs(l1[:, 0] - l0[:, 0]).max())
    assert np.allclose(l1, l0, rtol=5e-4, atol=1e-5), float(np.abs(l1 - l0).max())
    d = np.abs(p1 - p0)
    assert d.max() <= steps * 2e-4 * 1.4 * 1.01                      # lr_steps peak = 1.3e-4; one sign flip per step at most
    assert np.quantile(d, 0.995) < 1e-4
