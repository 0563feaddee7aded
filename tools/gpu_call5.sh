#!/bin/bash
mkdir -p gpurun_out
for g in 37 74 111 148; do
NAB200_MAX_GRID_CTAS=$g timeout 600 python bench.py --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/bench_g$g.json 2>/dev/null; cat gpurun_out/bench_g$g.json | python -c "import sys,json; d=json.load(sys.stdin); print('grid ctas/SM', $g/37, 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
