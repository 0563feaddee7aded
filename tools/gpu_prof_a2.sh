#!/bin/bash
# instruction profile of one stream pass (the prewarm launches run one stream) and a full-batch capture of the A2 kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 10 -c 1 -f -o gpurun_out/prof_a2_one python tools/h_check.py timing_a2 > gpurun_out/ncu_a2.out 2>&1; tail -2 gpurun_out/ncu_a2.out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 80 -c 1 -f -o gpurun_out/prof_a2 python tools/h_check.py timing_a2 > gpurun_out/ncu_a2b.out 2>&1; tail -2 gpurun_out/ncu_a2b.out
