#!/bin/bash
# same-box A/B of the fp32 LSTM kernels' packed activation (library builds with libna_t.so = -DNAB_LSTM_ACT=0, the IEEE quotient; default 2): bench line of cfg 3 + the other shapes
for lib in libna_t.so libneuralaudio_b200.so; do
NAB200_LIBNAME=$lib timeout 200 python bench.py --workload lstm_1x16 --steps 100 --no-cpu-baseline --no-extras --sustained-seconds 0 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('$lib lstm_1x16 8192x128 dev', round(d['ms_per_step']*1000,1), 'us', round(d['value']/1e9,2), 'Gs/s parity', d['parity']['max_abs'])"
NAB200_LIBNAME=$lib timeout 100 python tools/lstm_tc_check.py 2048 syn_lstm_2x8 syn_lstm_1x8 syn_lstm_1x24 syn_lstm_2x12 ref_BossLSTM_1x16 ref_BossLSTM_2x8 2>&1 | awk '{print "   ", $1, $2, $3, "other", $8, "us"}'
done
