# round 2, GPU call 6: step kernel variants (RNG hoist, row prefetch, table form, PDL) at d = 1
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --mode step --no-cpu-baseline --steps 20 > gpurun_out/r2c6_$tag.json 2> gpurun_out/r2c6_$tag.err; tail -2 gpurun_out/r2c6_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r2c6_$tag.json')); print('$tag', 'us/step', d['ms_per_step']*10, 'kernel_us', d['roofline']['kernel_us'], 'e2e', d['e2e']['value'])"; }
run hoist1_pref1 GJB_STEP_HOIST=1 GJB_STEP_PREFETCH=1
run hoist0_pref1 GJB_STEP_HOIST=0 GJB_STEP_PREFETCH=1
run hoist1_pref0 GJB_STEP_HOIST=1 GJB_STEP_PREFETCH=0
run hoist0_pref0 GJB_STEP_HOIST=0 GJB_STEP_PREFETCH=0
run hoist1_pref1_nopdl GJB_STEP_HOIST=1 GJB_STEP_PREFETCH=1 GJB_PDL=0
run hoist0_pref0_nopdl GJB_STEP_HOIST=0 GJB_STEP_PREFETCH=0 GJB_PDL=0
run table GJB_STEP_TABLE=1
