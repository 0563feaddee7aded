#!/bin/bash
mkdir -p gpurun_out
echo "== base"; ./tools/h_bench 4096 0 0 | tail -2 | tr '\n' ' '; echo
echo "== aliased state (L2 hits)"; NAB_H_ALIAS=1 ./tools/h_bench 4096 0 0 | tail -2 | tr '\n' ' '; echo
NAB_H_ALIAS=1 ./tools/h_timing 4096 0 0 > gpurun_out/h_timing_alias.txt
echo "== no windows"; ./tools/h_bench_nowin 4096 0 0 | tail -2 | tr '\n' ' '; echo
./tools/h_timing_nowin 4096 0 0 > gpurun_out/h_timing_nowin.txt
