#!/usr/bin/env python3
"""Drop-in for the reference program of the same name: k-fold training and evaluation of the end-to-end
supervised model (dual health / disease decoders + BatchNorm / Dropout classifier + contrastive hinge) with the reference's
flags, every fold trained in one fused launch on libnmb's B200 kernels.  See multi_modal_normative_modeling_b200/e2e.py."""
from multi_modal_normative_modeling_b200.e2e import cli_main

if __name__ == "__main__":
    cli_main()
