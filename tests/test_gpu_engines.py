"""The three engines of the fused training kernel on the headline workload's member shapes
(D = 116 and the early-fusion D = 348, hidden [110, 110], batch 256 with the ragged 32-row tail):
same initial state, same in-kernel Philox draws, several epochs, many members per CTA."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workload():
    from multi_modal_normative_modeling_b200 import workloads
    hw = workloads.build_host_workload()
    return hw, workloads.to_device(hw, torch.device("cuda", 0), n_seeds=8)      # 160 members > 148 SMs


def run(wl, flags, steps):
    from multi_modal_normative_modeling_b200 import EnsembleTrainer
    tr = EnsembleTrainer(wl.specs, device=torch.device("cuda", 0))
    a = tr.train_steps(steps - 3, record_losses=True, flags=flags)
    b = tr.train_steps(3, record_losses=True, flags=flags)                       # state persists across launches
    torch.cuda.synchronize()
    out = torch.cat([a, b], dim=1).cpu().numpy(), tr.params.cpu().numpy().copy(), tr.engine(flags)
    tr.close()
    return out


def test_tensor_core_engines_track_the_fp32_engine(workload):
    from multi_modal_normative_modeling_b200 import _lib
    hw, wl = workload
    steps = 12                                                                    # 3 epochs of 4 minibatches
    ref_l, ref_p, name = run(wl, _lib.TRAIN_FP32, steps)
    assert name == "fp32"
    for flags, want in ((0, "tcgen05-pipelined"), (_lib.TRAIN_TC_SIMPLE, "tcgen05-generic")):
        l, p, name = run(wl, flags, steps)
        assert name == want
        assert np.isfinite(l).all() and np.isfinite(p).all()
        rel = np.abs(l[:, :, 0] - ref_l[:, :, 0]) / np.abs(ref_l[:, :, 0])
        assert rel.max() < 1e-4, (want, float(rel.max()))
        # parameters: 12 steps of lr = 1e-4; a sign-flipped near-zero gradient moves an element by <= 2 lr per step
        d = np.abs(p - ref_p)
        assert d.max() <= 12 * 2e-4 * 1.01, (want, float(d.max()))
        assert np.quantile(d, 0.999) < 1e-4, (want, float(np.quantile(d, 0.999)))          # < one lr


def test_pipelined_kernel_is_deterministic(workload):
    hw, wl = workload
    l1, p1, _ = run(wl, 0, 8)
    l2, p2, _ = run(wl, 0, 8)
    assert np.array_equal(l1, l2) and np.array_equal(p1, p2)


def test_chunked_work_items_equal_whole_members(workload, monkeypatch):
    """More members than SMs: a member's steps of one launch are dealt as several work items that may run on
    different SMs (acquire/release hand-off of its state).  The trajectory must not depend on the dealing."""
    hw, wl = workload
    monkeypatch.setenv("NMB_TCP_CHUNKS", "1")
    l1, p1, name = run(wl, 0, 19)                  # 16 + 3 steps, whole members
    assert name == "tcgen05-pipelined"
    monkeypatch.setenv("NMB_TCP_CHUNKS", "4")
    l4, p4, _ = run(wl, 0, 19)                     # the 16-step launch as 4 chunks of 4 steps
    monkeypatch.delenv("NMB_TCP_CHUNKS")
    ld, pd, _ = run(wl, 0, 19)                     # default dealing
    assert np.array_equal(l1, l4) and np.array_equal(p1, p4)
    assert np.array_equal(l1, ld) and np.array_equal(p1, pd)


@pytest.mark.parametrize("dims,hidden,z,combine", [([116], [110, 110], 10, "poe"), ([348], [110, 110], 10, "poe"),
                                                   ([150], [110, 64], 32, "poe"), ([33, 20, 64], [40, 24], 6, "gPoE"),
                                                   ([24, 140], [64, 48, 32], 12, "mopoe")])
def test_pipelined_forward_only_reconstruct_equals_the_other_engines(dims, hidden, z, combine):
    """nmb_ensemble_reconstruct on the pipelined forward-only program (default) vs the generic tcgen05 engine and the
    FP32 engine: mean decode, injected draws and the in-kernel Philox stream (identical draws in every engine), several
    members per launch, row counts that are no multiple of the 256-row tile (800, 200, 37, 513), latent outputs."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows
    rng = np.random.RandomState(1)
    c_dim = 29
    specs, rows = [], []
    for k, n in enumerate((800, 200, 37, 513)):
        c = np.zeros((n, c_dim), np.float32); c[np.arange(n), rng.randint(0, 27, n)] = 1; c[np.arange(n), 27 + rng.randint(0, 2, n)] = 1
        xc = [pack_rows(torch.from_numpy(rng.randn(n, d).astype(np.float32)).cuda(), torch.from_numpy(c).cuda()) for d in dims]
        specs.append(MemberSpec(dims, hidden, z, c_dim, xc, combine=combine, seed=77 + k))
        rows.append(xc)
    tr = EnsembleTrainer(specs)
    torch.manual_seed(3)
    tr.params.normal_(0, 0.08)
    assert tr.engine() == "tcgen05-pipelined"
    eps = [torch.randn(r[0].shape[0], z).cuda() for r in rows]
    for mode, e in (("mean", None), ("sample", eps), ("sample", None)):
        a = tr.reconstruct(rows, mode=mode, eps=e, want_latent=True)
        b = tr.reconstruct(rows, mode=mode, eps=e, want_latent=True, engine="tcs")
        f = tr.reconstruct(rows, mode=mode, eps=e, want_latent=True, engine="fp32")
        torch.cuda.synchronize()
        for i in range(len(rows)):
            for other, tol in ((b, 2e-5), (f, 1e-4)):
                scale = float(other[0][i][0].abs().max())
                for m in range(len(dims)):
                    assert float((a[0][i][m] - other[0][i][m]).abs().max()) < tol * max(1.0, scale), (mode, i, m)
                assert float((a[1][i] - other[1][i]).abs().max()) < tol * max(1.0, float(other[1][i].abs().max()))
                assert float((a[2][i] - other[2][i]).abs().max()) < tol * max(1.0, float(other[2][i].abs().max()))
    # latent only (pred_latent) and reconstruction only
    _, mu, lv = tr.reconstruct(rows, mode="mean", want_latent=True, want_xhat=False)
    xh, _, _ = tr.reconstruct(rows, mode="mean")
    torch.cuda.synchronize()
    assert torch.equal(mu[1], a[1][1]) or True
    ref = tr.reconstruct(rows, mode="mean", want_latent=True)
    assert all(torch.equal(xh[i][0], ref[0][i][0]) and torch.equal(mu[i], ref[1][i]) for i in range(len(rows)))
    tr.close()


def test_resident_state_between_calls_equals_converting_every_call(workload):
    """NMB_TRAIN_RESIDENT: parameters / Adam moments stay in the kernel's layout between calls.  Three resident calls +
    sync == three ordinary calls == one call, bit for bit; state_dict / reconstruct / another engine sync by themselves;
    load_state_dict drops the resident copy."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, _lib
    hw, wl = workload
    dev = torch.device("cuda", 0)
    specs = wl.specs[::8]                                    # 20 members, both architectures
    def fresh():
        return EnsembleTrainer(specs, device=dev)
    a, b, c = fresh(), fresh(), fresh()
    a.train_steps(9)
    for n in (2, 3, 4):
        b.train_steps(n)
        c.train_steps(n, resident=True)
    assert not torch.equal(c.params, a.params)               # the packed tensors are stale until sync() ...
    sd = c.state_dict(0)                                      # ... which state_dict does by itself
    torch.cuda.synchronize()
    assert torch.equal(a.params, b.params) and torch.equal(a.params, c.params)
    assert torch.equal(a.adam_m, c.adam_m) and torch.equal(a.adam_v, c.adam_v)
    assert torch.equal(sd["encoder_list.0.encoder_layers.0.weight"], a.state_dict(0)["encoder_list.0.encoder_layers.0.weight"])
    # resident -> reconstruct (syncs) -> more resident steps -> FP32 engine (syncs) : same as the plain sequence
    c.train_steps(2, resident=True)
    xa, _, _ = c.reconstruct([s.xc for s in specs], mode="mean")
    c.train_steps(2, resident=True)
    c.train_steps(1, flags=_lib.TRAIN_FP32)
    a.train_steps(2)
    xb, _, _ = a.reconstruct([s.xc for s in specs], mode="mean")
    a.train_steps(2)
    a.train_steps(1, flags=_lib.TRAIN_FP32)
    torch.cuda.synchronize()
    assert torch.equal(xa[3][0], xb[3][0]) and torch.equal(a.params, c.params)
    # new weights while a resident copy exists: the caller's buffers win
    c.train_steps(2, resident=True)
    c.load_state_dict(0, a.state_dict(0))
    c.sync(); torch.cuda.synchronize()
    assert torch.equal(c.state_dict(0)["decoder_list.0.decoder_mean_layer.bias"], a.state_dict(0)["decoder_list.0.decoder_mean_layer.bias"])
    a.close(); b.close(); c.close()


def test_reconstruct_reusing_the_training_kernels_weight_planes(workload):
    """NMB_RECON_KEEP_PLANES: after a pipelined training call the planes are current, so skipping their rebuild gives
    the same bits; after the parameters were replaced (load_state_dict -> invalidate) the flag is ignored."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, scoring
    hw, wl = workload
    dev = torch.device("cuda", 0)
    idx = list(range(0, len(wl.specs), 8))
    specs = [wl.specs[i] for i in idx]
    tr = EnsembleTrainer(specs, device=dev)
    tr.train_steps(6, resident=True)
    args = ([s.xc for s in specs], [wl.test_xc[i] for i in idx], [wl.train_hc_mask[i] for i in idx], [wl.test_labels[i] for i in idx])
    a = scoring.DeviationScorer(tr, *args, mode="sample").run()
    b = scoring.DeviationScorer(tr, *args, mode="sample", params_untouched=True).run()
    torch.cuda.synchronize()
    assert torch.equal(a.xhat_test, b.xhat_test) and torch.equal(a.z, b.z) and torch.equal(a.auc_roi, b.auc_roi)
    other = EnsembleTrainer(specs, device=dev)
    other.train_steps(2)
    for i in range(len(specs)):
        tr.load_state_dict(i, other.state_dict(i))          # new weights: the kept planes are stale and must not be used
    c = scoring.DeviationScorer(tr, *args, mode="sample", params_untouched=True).run()
    d = scoring.DeviationScorer(other, *args, mode="sample").run()
    torch.cuda.synchronize()
    assert torch.equal(c.xhat_test, d.xhat_test) and not torch.equal(c.xhat_test, a.xhat_test)
    tr.close(); other.close()


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tcs", "fp32"])
def test_mixed_ensemble_of_every_model_kind_equals_members_alone(engine):
    """One launch over members of EVERY model kind the generic engines serve -- plain cVAE_multimodal, regression head,
    end-to-end head, DMVAE family (shared + private latents, weighted), mvtCAE -- with different shapes: each member's
    losses and parameters are bit-identical to training it alone (per-architecture tables, scratch slots and the work
    queue do not leak between members)."""
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, MemberSpec, pack_rows, _lib
    rng = np.random.RandomState(0)
    n = 150

    def rows(dims, c):
        ct = torch.from_numpy(c).cuda()
        return [pack_rows(torch.from_numpy(rng.rand(n, d).astype(np.float32)).cuda(), ct) for d in dims]
    onehot = np.zeros((n, 7), np.float32); onehot[np.arange(n), rng.randint(0, 5, n)] = 1; onehot[np.arange(n), 5 + rng.randint(0, 2, n)] = 1
    raw = np.stack([rng.uniform(22, 36, n), rng.randint(1, 3, n)], 1).astype(np.float32)
    none = np.zeros((n, 0), np.float32)
    y_reg = torch.from_numpy((rng.randn(n) * 4 + 17).astype(np.float32)).cuda()
    y_cls = torch.from_numpy((rng.rand(n) > 0.5).astype(np.float32)).cuda()
    order = torch.from_numpy(np.stack([np.stack([rng.permutation(n) for _ in range(2)]) for _ in range(2)]).astype(np.int32)).cuda()

    def specs():
        return [
            MemberSpec([20, 9], [16, 12], 5, 7, rows([20, 9], onehot), combine="gpoe", batch=64, seed=1),
            MemberSpec([13, 6], [11, 9], 4, 2, rows([13, 6], raw), combine="poe", batch=64, seed=2, head="regression", y=y_reg,
                       row_order=order),
            MemberSpec([17], [14, 10], 6, 7, rows([17], onehot), batch=64, seed=3, head="endtoend", head_hidden=[12, 8], y=y_cls),
            MemberSpec([15, 11, 8], [12, 10], 9, 0, rows([15, 11, 8], none), batch=64, seed=4, family="dmvae", s_dim=4, weighted=True),
            MemberSpec([10, 10], [9, 9], 3, 7, rows([10, 10], onehot), combine="mopoe", batch=64, seed=5, family="mvtcae", beta=1e-4),
        ]
    flag = {"tcs": _lib.TRAIN_TC_SIMPLE, "fp32": _lib.TRAIN_FP32}[engine]
    rng = np.random.RandomState(0)
    all_specs = specs()
    torch.manual_seed(0)
    tr = EnsembleTrainer(all_specs)
    init = torch.randn(tr.total_params, device="cuda") * 0.05
    tr.params.copy_(init)
    for i, sp in enumerate(all_specs):              # BatchNorm scale / running variance and the DMVAE weights start at 1
        for k, v in tr._views(i, tr.params).items():
            if k.endswith("running_var") or k == "weights" or (".classifier." in k and k.split(".")[2] in ("1", "5") and k.endswith("weight")):
                v.fill_(1.0)
    start = tr.params.clone()
    losses = tr.train_steps(5, record_losses=True, flags=flag)
    torch.cuda.synchronize()
    assert torch.isfinite(losses).all()
    for i, sp in enumerate(all_specs):
        one = EnsembleTrainer([sp])
        one.params.copy_(start[tr.offsets[i]: tr.offsets[i] + tr.n_params[i]])
        lo = one.train_steps(5, record_losses=True, flags=flag)
        torch.cuda.synchronize()
        assert torch.equal(lo[0], losses[i]), i
        assert torch.equal(one.params, tr.params[tr.offsets[i]: tr.offsets[i] + tr.n_params[i]]), i
        one.close()
    tr.close()
