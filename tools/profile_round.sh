# Round profile: bench line, scoring timing, ncu launch list of the libnmb kernels of a bench run, one full ncu
# capture of the training kernel (run under gpurun on one B200; outputs in gpurun_out/).
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 300 gpurun_out/bench_r01b.err
python tools/time_scoring.py > gpurun_out/scoring_r01b.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tcp_|xprep|recon|auc_kernel|deviation_kernel|stats_kernel|pack_rows" --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_tcp -s 3 -c 1 -o gpurun_out/r01b_tcp -f python tools/run_tcp.py 24 20 5 > gpurun_out/ncu_tcp_b.log 2>&1
python tools/trace_tcp.py 1 33 62 > gpurun_out/trace_r01b_d116.txt 2>&1
python tools/trace_tcp.py 1 46 98 big > gpurun_out/trace_r01b_d348.txt 2>&1
