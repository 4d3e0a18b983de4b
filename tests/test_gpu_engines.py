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
