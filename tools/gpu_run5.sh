set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r2e_tests.log 2>&1
bash profiles/run_ncu_r2.sh r2e
tail -4 gpurun_out/r2e_tests.log
