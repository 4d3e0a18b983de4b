"""GPU unit tests of the tcgen05 GEMM engine behind the fused kernels (nmb_debug_tc_gemm):
UMMA shared-memory / instruction descriptors for every operand-major combination, ragged M/N/K,
and the BF16x3 split's accuracy (|err| <= 3e-5 of the max |C|, vs fp64)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(m, n, k, a_kmajor, b_kmajor, seed=0, scale=1.0):
    from multi_modal_normative_modeling_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(seed)
    a = (rng.randn(m, k) * scale).astype(np.float32)
    b = rng.randn(n, k).astype(np.float32)
    r4 = lambda v: (v + 3) // 4 * 4

    def dev(mat, kmajor):
        i, kk = mat.shape
        if kmajor:
            buf = np.zeros((i, r4(kk)), np.float32); buf[:, :kk] = mat
        else:
            buf = np.zeros((kk, r4(i)), np.float32); buf[:, :i] = mat.T
        buf[buf == 0] = 0
        return torch.from_numpy(buf).cuda(), buf.shape[1]
    da, lda = dev(a, a_kmajor)
    db, ldb = dev(b, b_kmajor)
    ldc = r4(n)
    c = torch.full((m, ldc), float("nan"), device="cuda")
    _lib.check(lib.nmb_debug_tc_gemm(da.data_ptr(), lda, int(a_kmajor), db.data_ptr(), ldb, int(b_kmajor),
                                     c.data_ptr(), ldc, m, n, k, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = a.astype(np.float64) @ b.astype(np.float64).T
    got = c.cpu().numpy()[:, :n].astype(np.float64)
    return np.abs(got - want).max() / np.abs(want).max(), got, want


@pytest.mark.parametrize("a_kmajor,b_kmajor", [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize("m,n,k", [(128, 16, 16), (256, 112, 64), (256, 110, 146), (32, 20, 111), (200, 116, 111),
                                   (110, 146, 256), (8, 4, 7), (256, 348, 111), (300, 130, 200)])
def test_tc_gemm_layouts(m, n, k, a_kmajor, b_kmajor):
    err, got, want = run(m, n, k, a_kmajor, b_kmajor)
    assert np.isfinite(got).all()
    assert err < 3e-5, (err, m, n, k, a_kmajor, b_kmajor)


def test_tc_gemm_small_magnitudes():
    """Gradient-sized operands (1e-6) must not lose their low part to bf16 underflow."""
    err, _, _ = run(256, 110, 116, 1, 0, seed=3, scale=1e-6)
    assert err < 3e-5
