# round 2, GPU call 29 (final, 1 GPU): full -m gpu suite, smoke, default bench line, d = 32 line, reference arm, the scoring-only streaming
# shape, ncu launch list + one --set full capture of the step kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|SKIPPED|passed|failed|Error|assert" > gpurun_out/r2c29_gpu_tests.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2c29_gpu_tests.log | tail -12 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c29_$1.json') if l.startswith('{')][-1]); print('$1'.ljust(14), 'us/step %.2f kernel_us %.2f value %.3e e2e %.3e frac %.3f | %s' % (d['ms_per_step']*10/(d['config']['T']/100), d['roofline']['kernel_us'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['logZ_check'][:36]))"; }
timeout 400 python bench.py > gpurun_out/r2c29_bench_d1.json 2> gpurun_out/r2c29_bench_d1.err; tail -2 gpurun_out/r2c29_bench_d1.err | cut -c1-200; show bench_d1
timeout 300 python bench.py --dim 32 --no-cpu-baseline --steps 10 > gpurun_out/r2c29_bench_d32.json 2> gpurun_out/r2c29_bench_d32.err; tail -1 gpurun_out/r2c29_bench_d32.err | cut -c1-200; show bench_d32
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" | tee gpurun_out/r2c29_bench_ref.json | cut -c1-400
timeout 300 python scripts/bench_assess.py --dim 32 2>&1 | grep "^{" | tee gpurun_out/r2c29_assess_d32.json
timeout 300 python scripts/bench_assess.py --dim 8 --particles 16777216 2>&1 | grep "^{" | tee gpurun_out/r2c29_assess_d8.json
timeout 300 python scripts/bench_assess.py --dim 1 --particles 33554432 2>&1 | grep "^{" | tee gpurun_out/r2c29_assess_d1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2c29_launches_d1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 3 -c 1 -f -o gpurun_out/r2c29_prof_pf_step_kernel_d1 python scripts/profile_pf.py --dim 1 --T 10 --mode step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size --clock-control none -k regex:model_kernel -c 3 --csv --log-file gpurun_out/r2c29_ncu_assess_d32.csv python scripts/bench_assess.py --dim 32 --iters 1 --warmup 0 > gpurun_out/ncu_assess.log 2>&1; tail -2 gpurun_out/ncu_assess.log | cut -c1-300
