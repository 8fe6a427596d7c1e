set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests/test_dist.py -m gpu -q -k "two_real" 2>&1 | tail -15) > gpurun_out/r2g_tests2.log 2>&1
tail -5 gpurun_out/r2g_tests2.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err)
tail -c 3000 gpurun_out/r2g_bench2.json; tail -12 gpurun_out/r2g_bench2.err
(P3_MG_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-verify > gpurun_out/r2g_bench2_nccl.json 2> gpurun_out/r2g_bench2_nccl.err)
tail -c 1500 gpurun_out/r2g_bench2_nccl.json; tail -5 gpurun_out/r2g_bench2_nccl.err
