set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dist.py tests/test_host_side.py -m gpu -q 2>&1 | tail -6) > gpurun_out/r2i_tests.log 2>&1
tail -3 gpurun_out/r2i_tests.log
(timeout 600 python bench.py --config 0 --steps 2 --warmup 1 > gpurun_out/r2i_config0.json 2> gpurun_out/r2i_config0.err); echo rc0=$?
tail -c 1500 gpurun_out/r2i_config0.json; tail -3 gpurun_out/r2i_config0.err
(timeout 900 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r2i_config4.json 2> gpurun_out/r2i_config4.err); echo rc4=$?
tail -c 600 gpurun_out/r2i_config4.json; tail -3 gpurun_out/r2i_config4.err
(timeout 900 python bench.py --config 2 --steps 1 --warmup 1 > gpurun_out/r2i_config2.json 2> gpurun_out/r2i_config2.err); echo rc2=$?
tail -c 1500 gpurun_out/r2i_config2.json; tail -3 gpurun_out/r2i_config2.err
