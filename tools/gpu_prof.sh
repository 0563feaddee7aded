#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavenet_ts -s 40 -c 1 -f -o gpurun_out/prof_ts python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.out 2>&1; tail -2 gpurun_out/ncu_full.out
