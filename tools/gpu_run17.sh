set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -12) > gpurun_out/r2q_tests.log 2>&1
tail -4 gpurun_out/r2q_tests.log
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err)
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json'))
print(d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()}, d['e2e']['ms_per_step'], d['expected_counts'] is not None, d['verified'], d['hbm_peak_bytes']/1e9, d['roofline']['frac'], d['cpu_baseline'])"; tail -3 gpurun_out/r2q_bench.err
(timeout 300 python bench.py --config 0 --steps 2 --warmup 1 > gpurun_out/r02_config0.json 2> gpurun_out/r2q_c0.err); tail -2 gpurun_out/r2q_c0.err; cut -c1-600 gpurun_out/r02_config0.json
(timeout 400 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r02_config4.json 2> gpurun_out/r2q_c4.err); tail -12 gpurun_out/r2q_c4.err
(timeout 400 python bench.py --config 2 --steps 1 --warmup 1 > gpurun_out/r02_config2.json 2> gpurun_out/r2q_c2.err); tail -4 gpurun_out/r2q_c2.err
