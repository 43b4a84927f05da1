set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d1.json 2> gpurun_out/bench_d1.err; tail -3 gpurun_out/bench_d1.err; cat gpurun_out/bench_d1.json
python bench.py --steps 5 --warmup 3 --dim 32 --no-cpu-baseline > gpurun_out/bench_d32.json 2> gpurun_out/bench_d32.err; tail -3 gpurun_out/bench_d32.err; cat gpurun_out/bench_d32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_d1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d1.log 2>&1
tail -2 gpurun_out/ncu_d1.log
