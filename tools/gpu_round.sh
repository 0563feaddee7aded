#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the dominant kernels.  Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/parity_report.py > gpurun_out/parity.txt 2>&1; tail -3 gpurun_out/parity.txt
timeout 300 python tools/latency_probe.py > gpurun_out/latency.txt 2>&1; cat gpurun_out/latency.txt
timeout 300 ./tools/model_test -s 4096 --kat oracle/_ref/models/BossWN-standard.nam oracle/_ref/models/BossWN-nano.nam oracle/_ref/models/BossLSTM-1x16.nam oracle/_ref/models/BossWN-a2.nam > gpurun_out/model_test.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_a1_standard.json 2> gpurun_out/bench_a1_standard.err; cat gpurun_out/bench_a1_standard.json; tail -3 gpurun_out/bench_a1_standard.err
timeout 300 python bench.py --workload lstm_1x16 --steps 50 --warmup 5 --cpu-seconds 4 > gpurun_out/bench_lstm_1x16.json 2>/dev/null; cat gpurun_out/bench_lstm_1x16.json
timeout 300 python bench.py --workload a2_full --steps 30 --warmup 5 --cpu-seconds 4 > gpurun_out/bench_a2_full.json 2>/dev/null; cat gpurun_out/bench_a2_full.json
timeout 300 python bench.py --workload a1_nano --steps 50 --warmup 5 --cpu-seconds 4 > gpurun_out/bench_a1_nano.json 2>/dev/null; cat gpurun_out/bench_a1_nano.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
# every launch with its device time (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.out 2>&1
# full capture of the dominant kernel: skip the S = 1 launches of the prewarm settle pass (receptive field / 128 + 2 of them)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 36 -c 1 -f -o gpurun_out/prof_wavenet python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd -s 20 -c 1 -f -o gpurun_out/prof_lstm python bench.py --workload lstm_1x16 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_lstm.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_h -s 60 -c 1 -f -o gpurun_out/prof_a2 python bench.py --workload a2_full --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_a2.out 2>&1
# the tensor-core LSTM kernel: 2x16 at 8192 streams is its automatic choice (skip the 2048-step prewarm launch)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 1 -c 1 -f -o gpurun_out/prof_lstm_tc python tools/lstm_tc_check.py 8192 syn_lstm_2x16 > gpurun_out/ncu_full_lstm_tc.out 2>&1
timeout 300 python tools/shape_table.py > gpurun_out/shape_table.txt 2>&1; tail -12 gpurun_out/shape_table.txt
(timeout 200 python tools/lstm_tc_check.py 8192; timeout 200 python tools/lstm_tc_check.py 32768 syn_lstm_1x16 syn_lstm_1x24 syn_lstm_2x8 syn_lstm_2x12 syn_lstm_2x16 syn_dyn_lstm_2x32) > gpurun_out/lstm_tc_check.txt 2>&1; cat gpurun_out/lstm_tc_check.txt
timeout 300 bash tools/gpu_blocking_ab.sh > gpurun_out/host_paths_ab.txt 2>&1; cat gpurun_out/host_paths_ab.txt
ls -la gpurun_out
