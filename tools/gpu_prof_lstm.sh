#!/bin/bash
# ncu --set full capture of one steady-state LSTM launch per library build in $@ (default: the product library)
mkdir -p gpurun_out
libs="$@"; [ -z "$libs" ] && libs=libneuralaudio_b200.so
for lib in $libs; do
  tag=${lib%.so}; tag=${tag#lib}
  NAB200_LIBNAME=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd -s 20 -c 1 -f -o gpurun_out/prof_lstm_$tag python bench.py --workload lstm_1x16 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm_$tag.out 2>&1; tail -1 gpurun_out/ncu_lstm_$tag.out
done
