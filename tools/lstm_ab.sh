mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lstm or LSTM or tw40" 2>&1 | tail -5
for r in 1 2; do for lib in libna_lstm0.so libneuralaudio_b200.so; do for w in lstm_1x16; do
NAB200_LIBNAME=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --steps 200 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('$lib $w', d['ms_per_step'], d['value']/1e9, d['clocks'])"
done; done; done
for lib in libna_lstm0.so libneuralaudio_b200.so; do NAB200_LIBNAME=$lib python tools/latency_probe.py oracle/_ref/models/BossLSTM-1x16.nam oracle/_ref/models/BossLSTM-2x8.nam; done
