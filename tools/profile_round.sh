#!/bin/sh
# Round profile (run under gpurun on one B200; outputs in gpurun_out/, turned into profiles/r02_* by tools/make_profiles.py):
# the bench line, the scoring breakdown, the ncu launch list of every libnmb kernel of a bench run, one full ncu capture
# of the training kernel and of the forward-only (reconstruction) instantiation, the per-item timelines.
set -x
P=r02
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; tail -c 300 gpurun_out/bench_$P.err
python tools/time_scoring.py > gpurun_out/scoring_$P.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tcp_|xprep|recon|auc_kernel|deviation_kernel|stats_kernel|pack_|robust|rank_bins|latent" --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-module-step > gpurun_out/ncu_bench_$P.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:train_tcp_kernelILb0 -s 3 -c 1 -o gpurun_out/${P}_tcp -f python tools/run_tcp.py 24 20 5 > gpurun_out/ncu_tcp_$P.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:train_tcp_kernelILb1 -s 4 -c 1 -o gpurun_out/${P}_recon -f python tools/time_scoring.py > gpurun_out/ncu_recon_$P.log 2>&1
python tools/trace_tcp.py 1 33 42 > gpurun_out/trace_${P}_d116.txt 2>&1
python tools/trace_tcp.py 1 46 84 big > gpurun_out/trace_${P}_d348.txt 2>&1
python tools/time_cli.py > gpurun_out/cli_$P.json 2> gpurun_out/cli_$P.err
python tests/tools/time_f3.py > gpurun_out/f3_$P.json 2> gpurun_out/f3_$P.err
ls -la gpurun_out/${P}_*
