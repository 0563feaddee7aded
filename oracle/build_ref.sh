#!/bin/bash
# Builds the UNMODIFIED reference Internal path (mikeoliphant/NeuralAudio) from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Nothing is copied from the reference tree except the model fixtures (CC BY-NC-ND, staged, never committed).
# Recipe follows SURVEY.md §8c: CI option set (static WaveNet/LSTM/A2), FastMath, 64-frame chunks.
# This is TEST INFRASTRUCTURE (the parity checker and the CPU baseline), never the product path.
set -e
R=${NA_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT/models"
if [ ! -d "$R/NeuralAudio" ]; then
  echo "build_ref: $R not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
CXX=/usr/bin/g++
FLAGS="-std=gnu++20 -O3  -w -fPIC \
 -DBUILD_INTERNAL_STATIC_WAVENET -DBUILD_INTERNAL_STATIC_LSTM -DBUILD_STATIC_INTERNAL_NAMA2 -DWAVENET_MAX_NUM_FRAMES=64 \
 -DLAYER_ARRAY_BUFFER_PADDING=24 -DWAVENET_MATH=FastMath -DLSTM_MATH=FastMath -DRTNEURAL_USE_EIGEN=1 -DRTNEURAL_DEFAULT_ALIGNMENT=16 \
 -DRTNEURAL_NAMESPACE=RTNeural -DDEFAULT_INPUT_DBU=12 -DDEFAULT_QUALITY_SCALE=1.0 \
 -I$R -I$R/NeuralAudio -isystem $R/deps/RTNeural -isystem $R/deps/math_approx/include \
 -isystem $R/deps/RTNeural/modules/Eigen -isystem $R/deps/RTNeural/modules/json"
SRCS="$R/NeuralAudio/NeuralModel.cpp $R/NeuralAudio/RTNeuralLoader.cpp $R/NeuralAudioCAPI/NeuralAudioCApi.cpp $R/deps/RTNeural/RTNeural/RTNeural.cpp"
STAMP="$OUT/.stamp"
NEW="$(cat "$0" "$HERE/ref_extra.cpp" "$HERE/ref_exports.map" | md5sum | cut -d' ' -f1)"
if [ -f "$OUT/libna_ref.so" ] && [ -f "$OUT/libna_ref_v4.so" ] && [ "$(cat "$STAMP" 2>/dev/null)" = "$NEW" ]; then
  echo "build_ref: up to date"
else
  # two ISA levels: x86-64-v3 (AVX2+FMA, runs on any GPU-box host) and x86-64-v4 (AVX-512, picked at
  # run time when /proc/cpuinfo has avx512f) so the CPU baseline is the reference at its best on that host.
  pids=""
  for lvl in v3 v4; do
    i=0
    for s in $SRCS "$HERE/ref_extra.cpp"; do
      $CXX $FLAGS -march=x86-64-$lvl -mtune=generic -c "$s" -o "$OUT/obj_${lvl}_$i.o" &
      pids="$pids $!"
      i=$((i+1))
    done
  done
  for p in $pids; do wait $p; done
  $CXX -shared -o "$OUT/libna_ref.so" "$OUT"/obj_v3_*.o -Wl,--version-script="$HERE/ref_exports.map" -Wl,-Bsymbolic -lpthread
  $CXX -shared -o "$OUT/libna_ref_v4.so" "$OUT"/obj_v4_*.o -Wl,--version-script="$HERE/ref_exports.map" -Wl,-Bsymbolic -lpthread
  rm -f "$OUT"/obj_*.o
  echo "$NEW" > "$STAMP"
fi
# A third build WITH the reference's NAM Core back-end (BUILD_NAMCORE, NeuralAudio/CMakeLists.txt:14-17,145-171): what the reference itself
# runs for the A2 files its Internal path refuses (NeuralModel.cpp:365-380), e.g. an A2 model on a host at 96 kHz.  Used only to
# generate golden vectors for those cases (tests/golden/make_golden.py --a2-oversampled-only with NA_REF_NAMCORE=1).
NAMC="$R/deps/NeuralAmpModelerCore"
if [ -d "$NAMC/NAM" ] && { [ ! -f "$OUT/libna_ref_namcore.so" ] || [ "$(cat "$STAMP.namcore" 2>/dev/null)" != "$NEW" ]; }; then
  pids=""; i=0
  for s in $SRCS "$HERE/ref_extra.cpp" $NAMC/NAM/activations.cpp $NAMC/NAM/conv1d.cpp $NAMC/NAM/get_dsp.cpp $NAMC/NAM/ring_buffer.cpp $NAMC/NAM/lstm.cpp \
           $NAMC/NAM/dsp.cpp $NAMC/NAM/container.cpp $NAMC/NAM/wavenet/slimmable.cpp $NAMC/NAM/wavenet/model.cpp $NAMC/NAM/wavenet/a2_fast.cpp; do
    $CXX $FLAGS -DBUILD_NAMCORE -DNAM_SAMPLE_FLOAT -isystem $NAMC -march=x86-64-v3 -mtune=generic -c "$s" -o "$OUT/obj_nc_$i.o" &
    pids="$pids $!"; i=$((i+1))
  done
  for p in $pids; do wait $p; done
  $CXX -shared -o "$OUT/libna_ref_namcore.so" "$OUT"/obj_nc_*.o -Wl,--version-script="$HERE/ref_exports.map" -Wl,-Bsymbolic -lpthread
  rm -f "$OUT"/obj_nc_*.o
  echo "$NEW" > "$STAMP.namcore"
fi
# stage fixtures (read in place from the reference, never committed)
cp -u "$R"/Utils/Models/*.nam "$R"/Utils/Models/*.json "$OUT/models/" 2>/dev/null || true
for f in wavenet.nam wavenet_a1_standard.nam lstm.nam; do
  [ -f "$R/deps/NeuralAmpModelerCore/example_models/$f" ] && cp -u "$R/deps/NeuralAmpModelerCore/example_models/$f" "$OUT/models/namcore_$f" || true
done
echo "build_ref: done -> $OUT/libna_ref.so"
