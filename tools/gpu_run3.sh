set -x
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -80) > gpurun_out/r2c_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err)
tail -40 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_bench.json | cut -c1-3000; tail -5 gpurun_out/r2c_bench.err
