#!/bin/bash
# profiles/run_ncu_r02.sh <tag> [full] — round-2 profile of ONE B200 (run under gpurun). For the SAME bench command:
#   gpurun_out/<tag>_plain.json    the un-profiled run (must exit 0 first)
#   gpurun_out/<tag>_launches.csv  every kernel launch of one warm-up + one timed step with its device time
#   gpurun_out/<tag>_full.ncu-rep  (with "full") ncu --set full of every kernel of the timed step
# 20 Mbp genome (same generator / coverage / error rate as configs[1]; tables still >> L2) so that ncu's
# ~40 replays per kernel stay within the time limit.
set -u
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --genome 20000000"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu1.log
if [ "${2:-}" = "full" ]; then
  N=$(grep -c '"gpu__time_duration.sum"' gpurun_out/${TAG}_launches.csv)
  HALF=$((N / 2))
  ncu --set full --clock-control none --import-source on -s $HALF -c $HALF \
      -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
  tail -n 2 gpurun_out/${TAG}_ncu2.log
fi
