# round 2, GPU call 19: all steps in one cooperative launch (mode="steps") vs one launch per step; full suite on the new build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c19_tests.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2c19_tests.log | tail -12 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c19_$1.json') if l.startswith('{')][-1]); print('$1'.ljust(14), 'us/step %.2f kernel_us %.2f value %.3e e2e %.3e frac %.3f | %s' % (d['ms_per_step']*10/(d['config']['T']/100), d['roofline']['kernel_us'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['logZ_check'][:36]))"; }
for m in step; do
  timeout 200 python bench.py --mode $m --no-cpu-baseline --steps 20 > gpurun_out/r2c19_d1_$m.json 2> gpurun_out/r2c19_d1_$m.err; tail -1 gpurun_out/r2c19_d1_$m.err | cut -c1-200; show d1_$m
  timeout 200 python bench.py --mode $m --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c19_d32_$m.json 2> gpurun_out/r2c19_d32_$m.err; tail -1 gpurun_out/r2c19_d32_$m.err | cut -c1-200; show d32_$m
done
timeout 400 python bench.py > gpurun_out/r2c19_bench_default.json 2> gpurun_out/r2c19_bench_default.err; tail -2 gpurun_out/r2c19_bench_default.err; cut -c1-200 gpurun_out/r2c19_bench_default.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2c19_launches_d1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
