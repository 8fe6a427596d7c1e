set -x
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r2u_bench.err)
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_1gpu.json'))
print(d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()}, d['e2e']['ms_per_step'], d['expected_counts'] is not None, d['verified'], d['hbm_peak_bytes']/1e9, d['roofline']['frac'])"; tail -3 gpurun_out/r2u_bench.err
bash profiles/run_ncu_r2.sh r02b
