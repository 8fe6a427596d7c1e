set -x
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -12) > gpurun_out/r2o_tests.log 2>&1
tail -4 gpurun_out/r2o_tests.log
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err)
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()}, d['e2e'], d['expected_counts'] is not None, d['verified'], d['hbm_peak_bytes']/1e9, d['roofline']['frac'], d['cpu_baseline'])"; tail -3 gpurun_out/r2o_bench.err
