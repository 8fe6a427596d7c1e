set -x
mkdir -p gpurun_out
(timeout 400 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r2j_config4.json 2> gpurun_out/r2j_config4.err); echo rc4=$?
tail -c 400 gpurun_out/r2j_config4.json; tail -12 gpurun_out/r2j_config4.err
(P3_SWEEP_K=63 P3_SWEEP_THR=2 timeout 120 python bench.py --config 4 --steps 1 --warmup 0 --genome 5000000 > gpurun_out/r2j_config4_k63_5M.json 2> gpurun_out/r2j_config4_k63_5M.err); echo rc463=$?
tail -c 600 gpurun_out/r2j_config4_k63_5M.json; tail -4 gpurun_out/r2j_config4_k63_5M.err
(P3_LONG_K=63 timeout 150 python bench.py --config 2 --steps 1 --warmup 1 --genome 5000000 > gpurun_out/r2j_config2_5M.json 2> gpurun_out/r2j_config2_5M.err); echo rc2=$?
tail -c 800 gpurun_out/r2j_config2_5M.json; tail -4 gpurun_out/r2j_config2_5M.err
timeout 200 ncu --set full --import-source on --clock-control none -k 'regex:insert_find' -s 1 -c 1 -o gpurun_out/r2j_find_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify --genome 20000000 > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_ncu.log | cut -c1-200
