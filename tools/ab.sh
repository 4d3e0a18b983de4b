#!/bin/sh
# Dev tool: A/B timing of library variants / knobs in ONE GPU session (same box, same clocks): cfg4 per-architecture and
# whole-ensemble step times of tools/time_subsets.py for each configuration.
set -x
for m in 1 3 5 7 1 7; do NMB_TCP_MERGE=$m python tools/time_subsets.py quick; done
