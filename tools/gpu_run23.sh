set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
(timeout 300 $T bench.py --config 3 --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_config3_8gpu_keys.json 2> gpurun_out/r2w_c3.err); tail -5 gpurun_out/r2w_c3.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_config3_8gpu_keys.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], d['lap_ms'], d['verified'], d['hbm_peak_bytes']/1e9, d['config']['insert_rounds'], d['config']['bloom_passes'], d['roofline']['frac'], d['counts'])"
