#!/bin/bash
# profiles/run_ncu_adj.sh <tag> — ncu --set full of adjacency_kernel at the full configs[1] size, once through
# the single-GPU path (bench.py) and once through the distributed algorithm emulated with one rank on one GPU
# (tools/mg_emulated_bench.py). Each command first runs without ncu.
set -u
TAG=$1
A="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
B="python tools/mg_emulated_bench.py --world 1 --genome 100000000 --steps 2"
$A > gpurun_out/${TAG}_single_plain.json 2> gpurun_out/${TAG}_single_plain.err &&
ncu --set full --clock-control none --import-source on -k "regex:adjacency_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_single -f $A > gpurun_out/${TAG}_single_ncu.log 2>&1
$B > gpurun_out/${TAG}_emul_plain.json 2> gpurun_out/${TAG}_emul_plain.err &&
ncu --set full --clock-control none --import-source on -k "regex:adjacency_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_emul -f $B > gpurun_out/${TAG}_emul_ncu.log 2>&1
tail -n 2 gpurun_out/${TAG}_single_ncu.log; tail -n 2 gpurun_out/${TAG}_emul_ncu.log
