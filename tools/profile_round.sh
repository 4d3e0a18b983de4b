set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 600 gpurun_out/bench_r01b.err
python tools/time_scoring.py > gpurun_out/scoring_r01b.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_tcp -s 3 -c 1 -o gpurun_out/r01b_tcp python tools/run_tcp.py 24 20 5 > gpurun_out/ncu_tcp_b.log 2>&1
ls -la gpurun_out | tail -5
