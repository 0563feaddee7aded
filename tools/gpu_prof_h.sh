#!/bin/bash
# parity + timing of the fp16-pair kernel, then one full ncu capture of it (5 streams per SM)
mkdir -p gpurun_out
NAB200_H_DEBUG=1 timeout 300 python tools/h_check.py parity 2>&1 | grep -v "^wavenet_h_kernel" | tail -3; NAB200_H_DEBUG=1 python tools/h_check.py timing1 2>&1 | tail -2
timeout 300 python tools/h_check.py timing 2>&1 | tail -6
NAB200_H_CTAS=${1:-5} timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 60 -c 1 -f -o gpurun_out/prof_h python tools/h_check.py timing1 > gpurun_out/ncu_h.out 2>&1; tail -2 gpurun_out/ncu_h.out
