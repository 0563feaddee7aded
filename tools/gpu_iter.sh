#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/ts_timing 4096 > gpurun_out/ts_timing_4096.txt 2>&1; head -6 gpurun_out/ts_timing_4096.txt
timeout 600 python tools/ts_check.py 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ts.json 2> gpurun_out/bench_ts.err; cat gpurun_out/bench_ts.json | python -c "import sys,json; d=json.load(sys.stdin); print('A1STD ms/step', d['ms_per_step'], 'Gs/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"; tail -3 gpurun_out/bench_ts.err
