#!/bin/bash
# same-box A/B of plain timing binaries: tools/gpu_ab.sh suffix1 suffix2 ... (h_bench<suffix>); A1 and A2 shapes, two rounds
mkdir -p gpurun_out
for round in 1 2; do
for sfx in "$@"; do
  [ "$sfx" = base ] && sfx=""
  a1=$(timeout 120 ./tools/h_bench$sfx 4096 0 0 | tail -1); a2=$(timeout 120 ./tools/h_bench$sfx 4096 1 0 | tail -1)
  echo "h_bench$sfx | A1 $a1 | A2 $a2"
done
done
