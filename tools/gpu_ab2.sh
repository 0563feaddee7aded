#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4; do echo "== R=512 ctas=$c"; timeout 120 ./tools/h_timing 4096 0 $c | sed -n 3,4p | tr '\n' ' '; echo; done
for c in 4 5 6; do echo "== R=320 (fake) ctas=$c"; NAB_H_FAKE_R=320 timeout 120 ./tools/h_timing 4096 0 $c > gpurun_out/h_fake_c$c.txt; sed -n 3,4p gpurun_out/h_fake_c$c.txt | tr '\n' ' '; echo; done
