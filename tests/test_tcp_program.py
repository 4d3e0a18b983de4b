"""CPU check of the host-built step program of the pipelined tcgen05 kernel (csrc/nmb_tcp.h): a small
simulator of the three kernel roles (TMA producer with a 3-slot ring, MMA issuer, two epilogue groups +
joint items) walks the program of several architectures and must drain it without a deadlock -- the
dependency indices, commit flags and item order are what the device code blindly trusts."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "tcp_program_sim.cu")
CSRC = os.path.join(ROOT, "multi_modal_normative_modeling_b200", "csrc")


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("sim") / "tcp_program_sim")
    subprocess.run([nvcc, "-std=c++17", "-I", CSRC, SRC, "-o", exe], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("args", [["116"], ["348"], ["116", "3"], ["20"], ["348", "2"], ["1000"], ["7", "4"],
                                  ["128"], ["129"], ["64", "2"], ["3"],
                                  # forward-only (reconstruction) programs of the same architectures
                                  ["116", "1", "fwd"], ["348", "1", "fwd"], ["116", "3", "fwd"], ["348", "2", "fwd"],
                                  ["1000", "1", "fwd"], ["7", "4", "fwd"], ["129", "1", "fwd"]])
def test_step_program_drains_without_deadlock(sim, args):
    r = subprocess.run([sim] + args, capture_output=True, text=True)
    assert r.returncode == 0 and "OK all done" in r.stdout, r.stdout + r.stderr
