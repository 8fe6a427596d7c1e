set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r2d_tests.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r2d_ncu.log 2>&1
tail -8 gpurun_out/r2d_tests.log; tail -3 gpurun_out/r2d_ncu.log | cut -c1-600; wc -l gpurun_out/r2d_launches.csv
