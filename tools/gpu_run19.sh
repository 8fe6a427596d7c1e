set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
(timeout 300 python -m pytest tests/test_dist.py -m gpu -q --maxfail=20 2>&1 | tail -6) > gpurun_out/r2s_dist.log 2>&1
tail -4 gpurun_out/r2s_dist.log
(timeout 400 $T bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r2s_b2.err); tail -3 gpurun_out/r2s_b2.err; cut -c1-200 gpurun_out/r02c_bench_2gpu.json
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02c_bench_2gpu.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['stage_ms'], d['lap_ms'], d['verified'], d['expected_counts'] is not None, d['hbm_peak_bytes']/1e9, d['e2e'])"
(timeout 500 $T bench.py --config 3 --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_config3_2gpu.json 2> gpurun_out/r2s_c3.err); tail -5 gpurun_out/r2s_c3.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_config3_2gpu.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], d['lap_ms'], d['verified'], d['hbm_peak_bytes']/1e9, d['config']['insert_rounds'], d['config']['bloom_passes'])"
(timeout 400 $T bench.py --config 2 --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_config2_2gpu.json 2> gpurun_out/r2s_c2.err); tail -5 gpurun_out/r2s_c2.err; cut -c1-300 gpurun_out/r02_config2_2gpu.json
