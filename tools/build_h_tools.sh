#!/bin/bash
# builds the fp16-pair kernel's stand-alone timing binaries: tools/h_timing (cycle stamps) and tools/h_bench (plain; for ncu)
# usage: tools/build_h_tools.sh [suffix] [extra -D flags...]
cd "$(dirname "$0")/.."
SUF=$1; shift
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DNAB_H_TOOLS -Iinclude -Ineuralaudio_b200/csrc"
$NV -DNAB_H_TIMING "$@" -o tools/h_timing$SUF tools/h_timing.cu neuralaudio_b200/csrc/model_desc.cpp -x cu &
$NV "$@" -o tools/h_bench$SUF tools/h_timing.cu neuralaudio_b200/csrc/model_desc.cpp -x cu &
wait
