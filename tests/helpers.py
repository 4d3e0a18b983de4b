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


def drop_knife_rows(g, key, got, want):
    """Rows of a hidden layer's weight / bias gradient whose unit sits on the leaky-relu knife edge in the
    reference's own step (|pre-activation| < 2e-6 of the layer's scale for some sample; recorded by
    oracle/make_golden.py as knife/<layer>): the reference's value there is decided by fp32 rounding noise."""
    layer = key.rsplit(".", 1)[0]
    units = g.get("knife/" + layer)
    if units is None:
        return got, want
    keep = np.ones(want.shape[0], dtype=bool)
    keep[units] = False
    return np.asarray(got)[keep], np.asarray(want)[keep]
