#!/bin/bash
# profiles/run_ncu_traffic.sh <tag> — DRAM bytes of every kernel of ONE step of the full configs[1]
# workload (two metrics only, so each kernel is replayed once or twice, not ~40 times).
set -u
TAG=${1:-r01e}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
KERN='regex:hist21|scatter21|insert_bins|cand_check|solid_kernel|makebf|compact_set|bloom_list|seeds_kernel|adjacency'
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$KERN" -s 17 -c 17 --csv \
    --log-file gpurun_out/${TAG}_traffic.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
