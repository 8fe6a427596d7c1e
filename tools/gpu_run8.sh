set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2h_bench8.json 2> gpurun_out/r2h_bench8.err)
echo rc=$?
tail -c 2500 gpurun_out/r2h_bench8.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2h_bench8.err | tail -15
