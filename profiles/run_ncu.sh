#!/bin/bash
# profiles/run_ncu.sh <tag> — run under gpurun on ONE B200. Produces, for the SAME bench command:
#   gpurun_out/<tag>_plain.json    the un-profiled run (must exit 0 first)
#   gpurun_out/<tag>_launches.csv  every launch of our kernels with its device time
#   gpurun_out/<tag>_full.ncu-rep  ncu --set full of the hot kernels of one timed step
# The profiled command uses a 20 Mbp genome (same generator/coverage/error rate as the headline
# config, tables still >> L2) so that ncu's ~40 replays per kernel stay within the time limit.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --genome 20000000"
KERN='regex:count21|flags21|hist21|scatter21|insert_bins|cand_check|solid_kernel|makebf|compact_set|bloom_list|seeds_kernel|adjacency|rend_kernel|export_counts|scan_parts'
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:hist21|scatter21|insert_bins|cand_check|makebf_kernel|bloom_list|adjacency' \
    -s 18 -c 18 -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
