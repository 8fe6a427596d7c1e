set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -30) > gpurun_out/r2v_tests.log 2>&1
tail -12 gpurun_out/r2v_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(timeout 400 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r02_config3_1gpu_keys.json 2> gpurun_out/r2v_c3.err); tail -4 gpurun_out/r2v_c3.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_config3_1gpu_keys.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], d['lap_ms'], d['verified'], d['hbm_peak_bytes']/1e9, d['config']['insert_rounds'], d['config']['bloom_passes'], d['counts'])"
(timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err); tail -2 gpurun_out/r2v_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['expected_counts'] is not None, d['verified'])"
