"""Shared helpers of the golden-vector tests (fixtures: tests/golden/*.npz, recorded from the unmodified
reference by oracle/make_golden.py)."""
import os

import numpy as np

# full-size cases recorded in round 2: N rows, batch with a partial last batch, `epochs` passes
LOOP_CASES = ["mm_M1_D150_full", "mm_M1_D348_full", "mm_M1_D1000_full", "mm_M4_full_gpoe", "mm_M4_full_poe",
              "mm_M4_full_moe", "mm_M4_full_mopoe", "mm_M1_Z32", "mm_M1_L4", "mm_M1_wide", "nmmlp_M3_full",
              "nmmlp_M2_small"]
GENERIC_ONLY = {"mm_M1_wide"}       # hidden width > 127: served by the generic tcgen05 engine


def load(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False))
    if "shared" in g:               # lean cases: inputs and initial weights live in another fixture (same seed)
        base = dict(np.load(os.path.join(golden_dir, str(g["shared"]) + ".npz"), allow_pickle=False))
        for k, v in base.items():
            if k.startswith("init/") or k == "c" or (k.startswith("x") and k[1:].isdigit()):
                g[k] = v
    return g


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def is_nmmlp(name):
    return name.startswith("nmmlp")


def loop_batches(n, b):
    return [(r0, min(b, n - r0)) for r0 in range(0, n, b)]


def knife_taint(g, key):
    """What the recorded knife-edge units (oracle/make_golden.py::clean_seed; normally none) do to the comparison of
    gradient tensor `key`: returns (skip_whole_tensor, rows_to_drop).  A unit whose pre-activation is within rounding
    distance of zero takes leaky-relu slope 1 or 0.01 depending on the summation order of the reference's own fp32
    dot product; that decides its own gradient row AND moves every gradient upstream of it by ~0.5 %."""
    rows, skip = None, False
    for k in g:
        if not k.startswith("knife/"):
            continue
        layer = k[len("knife/"):]                       # e.g. encoder_list.2.encoder_layers.0
        side, m, _, l = layer.split(".")
        l = int(l)
        if key.rsplit(".", 1)[0] == layer:
            rows = g[k]
        if side == "encoder_list":
            if key.startswith(f"encoder_list.{m}.encoder_layers.") and int(key.split(".")[3]) < l:
                skip = True
        else:                                           # a decoder unit taints everything before the decoder input
            if key.startswith("encoder_list.") or key.startswith("alpha_m_list."):
                skip = True
            if key.startswith(f"decoder_list.{m}.decoder_layers.") and int(key.split(".")[3]) < l:
                skip = True
    return skip, rows


def drop_knife_rows(g, key, got, want):
    skip, rows = knife_taint(g, key)
    got, want = np.asarray(got), np.asarray(want)
    if skip:
        return want[:0], want[:0]
    if rows is None:
        return got, want
    keep = np.ones(want.shape[0], dtype=bool)
    keep[rows] = False
    return got[keep], want[keep]


def assert_grads_close(g, prefix, grads, rel, to_numpy=lambda t: t):
    """Every gradient tensor recorded under `prefix` within `rel` of its max-norm; the gPoE alphas (a ParameterList of
    scalars) are compared as ONE vector.  Tensors absent from the reference must be exactly zero."""
    ref = sub(g, prefix)
    assert ref, prefix
    alphas = sorted(k for k in ref if k.startswith("alpha_m_list."))
    if alphas and not knife_taint(g, alphas[0])[0]:
        want = np.concatenate([ref[k].ravel() for k in alphas])
        got = np.concatenate([np.asarray(to_numpy(grads[k])).ravel() for k in alphas])
        assert np.abs(got - want).max() / (np.abs(want).max() + 1e-30) < rel, "alpha_m_list"
    for k, v in ref.items():
        if k in alphas:
            continue
        got, want = drop_knife_rows(g, k, np.asarray(to_numpy(grads[k])).reshape(v.shape), v)
        if want.size:
            assert np.abs(got - want).max() / (np.abs(v).max() + 1e-30) < rel, k
    for k in grads:
        if k not in ref:
            assert float(np.abs(np.asarray(to_numpy(grads[k]))).max()) == 0.0, k


def assert_losses_close(got, want, rel=1e-4):
    """Per-step (total, kl, ll).  total and ll at `rel`; the KL term alone is second-order small at initialisation
    (mu ~ logvar ~ 0), so once the weights have taken Adam steps -- which agree with the reference only to the sign
    ambiguity of near-zero gradients -- it is held to 10 x rel."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert np.allclose(got[..., 0], want[..., 0], rtol=rel), (got, want)
    assert np.allclose(got[..., 2], want[..., 2], rtol=rel), (got, want)
    assert np.allclose(got[..., 1], want[..., 1], rtol=10 * rel), (got, want)


def assert_update_close(name, got, final, init, steps, lr, exact_engine, g0=None, q99_tc=2e-3, mean_tc=1e-3):
    """Adam trajectory check on the UPDATE d = final - init (the parameter itself would pass trivially).

    Adam's normalised step lr * m_hat / sqrt(v_hat) is ~ lr * sign(.) in its first steps, so it is ill-conditioned
    wherever the running gradient average passes through zero: an element whose gradient is within rounding distance
    of zero -- or, over several steps, whose full-batch and ragged-batch gradients nearly cancel in m_hat -- may step
    the other way under ANY change of summation order.  Even oracle/cvae_torch.py (same fp32 ops as the reference, a
    different op order in the log-likelihood) shows this on up to 0.5 % of the elements of a tensor.  A real optimiser
    bug (bias correction, betas, lr, lost moments) moves EVERY element by >= 1 %.  Bounds, in units of max|d_ref|:
      * every element <= 2 (one sign flip per step) and |got - final| <= 2 * steps * lr;
      * 99 % of the elements < 2e-3 (FP32 FFMA engine: < 2e-4), mean deviation < 1e-3."""
    d_ref, d_got = final - init, got - init
    scale = np.abs(d_ref).max()
    if scale == 0:
        assert np.abs(d_got).max() == 0, name
        return
    dev = np.abs(d_got - d_ref) / scale
    assert dev.max() <= 2.0 + 1e-3, (name, float(dev.max()))
    assert np.abs(got - final).max() <= 2.0 * steps * lr * (1 + 1e-3), name
    flat = np.sort(dev.ravel())
    q99 = flat[min(flat.size - 1, int(np.ceil(0.99 * flat.size)))] if flat.size >= 200 else flat[max(0, flat.size - 3)]
    assert q99 < (2e-4 if exact_engine else q99_tc), (name, float(q99))
    if flat.size >= 200:
        assert dev.mean() < (1e-3 if exact_engine else mean_tc), (name, float(dev.mean()))
