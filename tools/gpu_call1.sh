#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 ./tools/mma_bench > gpurun_out/mma_bench.txt 2>&1; tail -5 gpurun_out/mma_bench.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json | python -c "import sys,json; d=json.load(sys.stdin); print('A1STD ms/step', d['ms_per_step'], 'Gs/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"; tail -3 gpurun_out/bench_quick.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavenet_tc -s 40 -c 1 -f -o gpurun_out/prof_wavenet python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.out 2>&1; tail -3 gpurun_out/ncu_full.out
ls -la gpurun_out
