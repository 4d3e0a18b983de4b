#!/bin/sh
# Dev tool: A/B timing of library variants / knobs in ONE GPU session (same box, same clocks).
set -x
L=multi_modal_normative_modeling_b200/lib
python tools/time_subsets.py quick
NMB_LIB=$L/libnmb_W.so python tools/time_subsets.py quick
python tools/time_subsets.py quick
NMB_LIB=$L/libnmb_W.so python tools/time_subsets.py quick
