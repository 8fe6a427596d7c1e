set -x
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-verify"
for v in default:0 p1000:1000 p520:520 p350:350; do
  name=${v%%:*}; parts=${v##*:}
  if [ "$parts" = "0" ]; then (timeout 200 $B > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err)
  else (P3_PARTS=$parts timeout 200 $B > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err); fi
  python -c "
import json; d=json.load(open('gpurun_out/r2k_$name.json')); print('$name', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()})"
done
(P3_SWEEP_K=63 P3_SWEEP_THR=2 timeout 300 python bench.py --config 4 --steps 1 --warmup 0 > gpurun_out/r2k_config4_k63.json 2> gpurun_out/r2k_config4_k63.err); echo rc463=$?
tail -c 700 gpurun_out/r2k_config4_k63.json; tail -3 gpurun_out/r2k_config4_k63.err
