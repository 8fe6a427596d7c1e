#!/bin/bash
# profiles/run_ncu2.sh <tag> <kernel-regex> [skip] [count] — ncu --set full on selected kernels of
# the 20 Mbp profiling workload (same command must first exit 0 without ncu).
set -u
TAG=$1; KERN=$2; SKIP=${3:-0}; CNT=${4:-4}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --genome 20000000"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k "regex:$KERN" -s $SKIP -c $CNT \
    -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
