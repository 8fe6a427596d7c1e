set -x
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2b_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err)
(timeout 300 ./profiles/microbench/insert_variants 24 1000000000 > gpurun_out/r2b_insert_variants.txt 2>&1)
tail -15 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_bench.json | cut -c1-3000; tail -5 gpurun_out/r2b_bench.err
