#!/bin/bash
# ncu full capture (with source) of the stand-alone fp16-pair kernel bench: tools/gpu_prof_hb.sh [binary] [a2flag] [ctas] [outname]
mkdir -p gpurun_out
B=${1:-h_bench}; A=${2:-0}; C=${3:-0}; O=${4:-prof_hb}
timeout 120 ./tools/$B 4096 $A $C | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 3 -c 1 -f -o gpurun_out/$O ./tools/$B 4096 $A $C > gpurun_out/$O.out 2>&1; tail -2 gpurun_out/$O.out
