set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q --maxfail=5 2>&1 | tail -5) > gpurun_out/r2p_tests.log 2>&1
tail -3 gpurun_out/r2p_tests.log
(timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err)
python -c "
import json; d=json.load(open('gpurun_out/r2p_bench.json'))
print(d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms'].items()}, {k:round(v,1) for k,v in d['count_substage'].items()}, d['e2e']['ms_per_step'], d['expected_counts'] is not None, d['verified'])"; tail -3 gpurun_out/r2p_bench.err
