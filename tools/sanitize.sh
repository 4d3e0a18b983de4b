#!/bin/sh
# compute-sanitizer evidence for the hand-written synchronisation of libnmb (run under gpurun on one B200):
# memcheck + racecheck + synccheck over a 160-member ensemble dealt as chunked work items (cross-SM acquire / release
# hand-off, bounded spins, proxy fences) followed by a scoring pass.  The spin watchdog is switched off
# (NMB_TCP_SPIN_BUDGET=0): the tools stretch time by orders of magnitude and a trap would poison the context.
# Summaries land in gpurun_out/sanitize_*.txt; copy them to profiles/.
export NMB_TCP_SPIN_BUDGET=0
export NMB_TCP_CHUNKS=4
OUT=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py 8 8 > $OUT/sanitize_$tool.txt 2>&1
  echo "exit $?" >> $OUT/sanitize_$tool.txt
  tail -6 $OUT/sanitize_$tool.txt
done
