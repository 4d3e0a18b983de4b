#!/usr/bin/env python3
"""Drop-in for the reference program of the same name (same flags, same files), running on
libnmb's fused B200 kernels.  See multi_modal_normative_modeling_b200/cli.py."""
import argparse

from multi_modal_normative_modeling_b200.cli import add_common_args, train_main

if __name__ == "__main__":
    train_main(add_common_args(argparse.ArgumentParser(), train=True).parse_args())
