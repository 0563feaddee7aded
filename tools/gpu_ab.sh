#!/bin/bash
# A/B of timing binaries: tools/gpu_ab.sh bin1 bin2 ... ; each run for A1 (ctas 4, 5) and A2
mkdir -p gpurun_out
for b in "$@"; do
  for c in 4 5; do
    echo "== $b A1 ctas=$c"; timeout 120 ./tools/$b 4096 0 $c > gpurun_out/${b}_a1_c$c.txt 2>&1; sed -n 2,5p gpurun_out/${b}_a1_c$c.txt | tr '\n' ' '; echo
  done
  echo "== $b A2"; timeout 120 ./tools/$b 4096 1 0 > gpurun_out/${b}_a2.txt 2>&1; sed -n 2,5p gpurun_out/${b}_a2.txt | tr '\n' ' '; echo
done
