#!/bin/sh
# Installs the UNMODIFIED reference module cVAE.py into baseline/_ref (git-ignored; it travels to the GPU box with the
# gpurun snapshot) so that `bench.py --impl reference` and the cpu_baseline leg time the reference's own classes
# (kind "reference") instead of the restated port.  The reference repository has no packaging, so a two-line setup.py
# is written next to a COPY of cVAE.py under /tmp (the reference tree is read-only); the installed file is byte-identical.
set -e
REF=${NMB_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
cp "$REF/cVAE.py" "$T/"
printf 'from setuptools import setup\nsetup(name="multi_modal_normative_modeling_reference", version="0", py_modules=["cVAE"])\n' > "$T/setup.py"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/baseline/_ref" "$T"
cmp "$HERE/baseline/_ref/cVAE.py" "$REF/cVAE.py" && echo "baseline/_ref/cVAE.py is byte-identical to the reference"
