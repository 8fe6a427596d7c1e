#!/bin/bash
# profiles/run_ncu_r2.sh <tag> — round-2 profile of ONE B200 (run under gpurun; every step bounded by `timeout`).
#   gpurun_out/<tag>_launches.csv   every launch of OUR kernels in one warm-up + one timed step of the full configs[1]
#                                   workload: device time, DRAM bytes, L2 hit rate, achieved occupancy, instructions
#   gpurun_out/<tag>_full.ncu-rep   ncu --set full of the dominant kernels on the 20 Mbp workload (short replays)
# The un-profiled run of the same command must exit 0 first.
set -u
TAG=${1:-r02}
FULLSZ="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify"
SMALL="$FULLSZ --genome 20000000"
OURS='regex:rend_kernel|init_cursors|check_cursors|scatter21|insert_find|insert_add|pos_bin|apply_bins|pos_clear|solid_kernel|scatter_kmer|set_sweep|makebf_kernel|compact_set|double_hash|bloom_bin|bloom_apply|bloom_list|seeds_kernel|adjacency|hist21|scan_parts|scatter_rec|scatter_pos|peer_sync|publish_counts'
timeout 300 $FULLSZ > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread \
    --clock-control none -k "$OURS" -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $FULLSZ > gpurun_out/${TAG}_ncu1.log 2>&1
tail -n 1 gpurun_out/${TAG}_ncu1.log | cut -c1-300
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:insert_find|insert_add|pos_bin|set_sweep|scatter_kmer|scatter21|adjacency' -s 9 -c 12 \
    -o gpurun_out/${TAG}_full -f $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
tail -n 1 gpurun_out/${TAG}_ncu2.log | cut -c1-300
ls -la gpurun_out/ | grep ${TAG}
