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
