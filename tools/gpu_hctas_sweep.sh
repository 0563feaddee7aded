#!/bin/bash
# fp16-pair kernel: streams in flight per SM (option h_ctas) over the two benched shapes
for k in 0 3 4 5; do for w in a1_standard a2_full; do
NAB200_H_CTAS=$k timeout 200 python bench.py --workload $w --steps 60 --no-cpu-baseline --no-extras --sustained-seconds 0 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('h_ctas=$k $w', round(d['ms_per_step']*1000,1), 'us', round(d['value']/1e9,3), 'Gs/s')"
done; done
