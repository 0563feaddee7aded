#!/bin/bash
# ncu --set full capture of one steady-state launch of a bench workload: tools/gpu_prof_wl.sh <workload> <kernel regex> [skip]
mkdir -p gpurun_out
wl=$1; rx=$2; skip=${3:-8}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_$wl python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$wl.out 2>&1; tail -1 gpurun_out/ncu_$wl.out
