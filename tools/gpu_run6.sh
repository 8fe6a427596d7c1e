set -x
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -30) > gpurun_out/r2f_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err)
tail -6 gpurun_out/r2f_tests.log; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['ms_per_step'], d['stage_ms'], d['count_substage'], d['e2e'], d['expected_counts'], d['verified'], d['hbm_peak_bytes']/1e9)"; tail -5 gpurun_out/r2f_bench.err
