#!/usr/bin/env python3
"""Drop-in for the reference program of the same name: k-fold training and evaluation of
``cVAE_multimodal_regression`` with the reference's flags (-R -H -C -P -E -K --batch_size -BaseLR), every fold trained in
one fused launch on libnmb's B200 kernels.  See multi_modal_normative_modeling_b200/regression.py."""
from multi_modal_normative_modeling_b200.regression import main

if __name__ == "__main__":
    main()
