#!/bin/bash
# compute-sanitizer over the kernels added last: tcgen05 LSTM (forced), wide run-time-shaped WaveNet, dilated A2 head
W="syn_dyn_48x24 syn_dyn_3arrays syn_a2_full_sr96000 syn_a2_lite_sr96000"
L="syn_lstm_1x16 syn_lstm_2x16 syn_lstm_1x24 syn_dyn_lstm_2x32"
for tool in memcheck racecheck synccheck; do
  echo "== $tool / run-time-shaped + oversampled A2"; timeout 280 compute-sanitizer --tool $tool python tools/sanitize_run.py $W 2>&1 | grep -E "ok|SUMMARY|Error|error|hazard" | head -12
  echo "== $tool / tcgen05 LSTM (lstm_kernel = 4)"; NAB200_LSTM_KERNEL=4 timeout 280 compute-sanitizer --tool $tool python tools/sanitize_run.py $L 2>&1 | grep -E "ok|SUMMARY|Error|error|hazard" | head -12
done
