#!/bin/bash
# profiles/run_ncu_p3.sh <tag> — round-1 (third pass) profile of ONE B200 (run under gpurun; every step bounded by `timeout`).
#   gpurun_out/<tag>_plain.json     the un-profiled 20 Mbp run (must exit 0 first)
#   gpurun_out/<tag>_launches.csv   every kernel launch (warm-up + timed + e2e steps) with its device time
#   gpurun_out/<tag>_new_full.ncu-rep  ncu --set full of the kernels that are new in round 2 (one timed step)
#   gpurun_out/<tag>_traffic.csv    DRAM bytes + device time of every kernel of ONE full-size configs[1] step
# The 20 Mbp genome (same generator / coverage / error rate as configs[1]; tables still >> L2) keeps ncu's ~40
# replays per kernel short. A first version that captured EVERY kernel with --set full ran into the 20-minute limit.
set -u
TAG=${1:-r01p3}
SMALL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --genome 20000000"
FULLSZ="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
OURS='regex:rend_kernel|init_cursors|scatter21|insert_bins|pos_bin|plane_clear|solid_kernel|scatter_kmer|set_sweep|makebf_kernel|compact_set|double_hash|bloom_bin|bloom_apply|bloom_list|seeds_kernel|adjacency|cand_check|hist21|scan_parts'
timeout 120 $SMALL > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
tail -n 1 gpurun_out/${TAG}_ncu1.log | cut -c1-200
timeout 200 ncu --set full --clock-control none -k 'regex:pos_bin|scatter_kmer|set_sweep|bloom_bin|bloom_apply|plane_clear' -s 16 -c 16 \
    -o gpurun_out/${TAG}_new_full -f $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
tail -n 1 gpurun_out/${TAG}_ncu2.log | cut -c1-200
timeout 100 $FULLSZ > gpurun_out/${TAG}_traffic_plain.json 2> gpurun_out/${TAG}_traffic_plain.err || exit 1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$OURS" -c 120 --csv \
    --log-file gpurun_out/${TAG}_traffic.csv $FULLSZ > gpurun_out/${TAG}_traffic_ncu.log 2>&1
tail -n 1 gpurun_out/${TAG}_traffic_ncu.log | cut -c1-200
ls -la gpurun_out/
