set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-verify"
run() { name=$1; shift; (env "$@" timeout 200 $B > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err); python -c "
import json; d=json.load(open('gpurun_out/r2l_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()})"; }
run mb48 P3_X=1
run mb72 P3_PART_MB=72
run mb100 P3_PART_MB=100
run b8 P3_FIND_BATCH=8
run b2 P3_FIND_BATCH=2
