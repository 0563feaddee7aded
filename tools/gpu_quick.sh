#!/bin/bash
# quick visit: parity tests + the headline bench line (+ optional extra command in $1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json | python -c "import sys,json; d=json.load(sys.stdin); print('A1STD ms/step', d['ms_per_step'], 'Gs/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"; tail -3 gpurun_out/bench_quick.err
if [ -n "$1" ]; then bash -c "$1"; fi
