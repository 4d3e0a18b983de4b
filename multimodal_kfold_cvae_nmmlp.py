#!/usr/bin/env python3
"""Drop-in for the reference program of the same name: positional ``action`` in {train, test, analyze, all} and the
flags of multimodal_kfold_cvae_nmmlp.py:646-660, running on libnmb's fused B200 kernels (healthy-control-only
training rows, -MSE reconstruction term, triangular cyclic learning rate).  See multi_modal_normative_modeling_b200/cli.py."""
from multi_modal_normative_modeling_b200.cli import nmmlp_main
# the reference pickles its script-local model class as __main__.cVAE_multimodal_endtoend: keep the name resolvable here
from multi_modal_normative_modeling_b200.cVAE import cVAE_multimodal_endtoend  # noqa: F401

if __name__ == "__main__":
    nmmlp_main()
