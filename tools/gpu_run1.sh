set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2a_tests.log 2>&1
(timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err)
(timeout 300 ./profiles/microbench/insert_variants 24 1000000000 > gpurun_out/r2a_insert_variants.txt 2>&1)
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json | cut -c1-1500; tail -3 gpurun_out/r2a_bench.err
