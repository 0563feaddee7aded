#!/bin/bash
# full-batch ncu capture of the fp16-pair kernel; usage: tools/gpu_prof_h.sh [a1|a2] [launches to skip]
mkdir -p gpurun_out
W=${1:-a1}
T=timing1; [ "$W" = a2 ] && T=timing_a2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s ${2:-50} -c 1 -f -o gpurun_out/prof_${W} python tools/h_check.py $T > gpurun_out/ncu_${W}.out 2>&1; tail -1 gpurun_out/ncu_${W}.out
