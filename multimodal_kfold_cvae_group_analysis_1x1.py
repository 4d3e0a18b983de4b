#!/usr/bin/env python3
"""Drop-in for the reference program of the same name (same flags, same files): ROC-AUC by the GPU
pair-counting kernel, Youden threshold / accuracy / sensitivity / specificity per fold.
See multi_modal_normative_modeling_b200/cli.py."""
import argparse

from multi_modal_normative_modeling_b200.cli import add_common_args, analysis_main

if __name__ == "__main__":
    analysis_main(add_common_args(argparse.ArgumentParser(), train=True).parse_args())
