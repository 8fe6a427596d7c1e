set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-verify"
run() { name=$1; shift; (env "$@" timeout 200 $B > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err); python -c "
import json; d=json.load(open('gpurun_out/r2m_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()})"; }
run b1 P3_FIND_BATCH=1
run b2 P3_FIND_BATCH=2
