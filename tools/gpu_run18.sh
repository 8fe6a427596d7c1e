set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dist.py -m gpu -q --maxfail=20 2>&1 | tail -40) > gpurun_out/r2r_dist.log 2>&1
tail -25 gpurun_out/r2r_dist.log
(timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_dist.py 2>&1 | tail -8) > gpurun_out/r2r_tests.log 2>&1
tail -4 gpurun_out/r2r_tests.log
(P3_CONFIG2_DIST=1 timeout 400 python bench.py --config 2 --steps 1 --warmup 1 > gpurun_out/r02_config2_dist_1gpu.json 2> gpurun_out/r2r_c2.err); tail -6 gpurun_out/r2r_c2.err; cut -c1-1500 gpurun_out/r02_config2_dist_1gpu.json
(timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r02_config3_1gpu.json 2> gpurun_out/r2r_c3.err); tail -8 gpurun_out/r2r_c3.err; cut -c1-3000 gpurun_out/r02_config3_1gpu.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
