#!/bin/bash
# one visit for the fp16-pair kernel: parity triage, stamped phase timing, plain timing.  usage: tools/gpu_h.sh [ctas list]
mkdir -p gpurun_out
timeout 600 python tools/h_check.py parity > gpurun_out/h_check.txt 2>&1; cat gpurun_out/h_check.txt | tail -14
for c in ${1:-0}; do
  echo "== h_bench A1 ctas=$c"; timeout 120 ./tools/h_bench 4096 0 $c | tail -2 | tr '\n' ' '; echo
done
echo "== h_bench A2"; timeout 120 ./tools/h_bench 4096 1 0 | tail -2 | tr '\n' ' '; echo
timeout 120 ./tools/h_timing 4096 0 0 > gpurun_out/h_timing_a1.txt 2>&1; head -5 gpurun_out/h_timing_a1.txt | tail -2
timeout 120 ./tools/h_timing 4096 1 0 > gpurun_out/h_timing_a2.txt 2>&1; head -5 gpurun_out/h_timing_a2.txt | tail -2
