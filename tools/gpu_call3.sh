#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ts_check.py 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ts.json 2> gpurun_out/bench_ts.err; cat gpurun_out/bench_ts.json | python -c "import sys,json; d=json.load(sys.stdin); print('A1STD ms/step', d['ms_per_step'], 'Gs/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"; tail -3 gpurun_out/bench_ts.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavenet_ts -s 40 -c 1 -f -o gpurun_out/prof_ts python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.out 2>&1; tail -2 gpurun_out/ncu_full.out
