#!/bin/bash
# same-box A/B of the host-buffer paths: staged (copy engines, slices) vs zero-copy (kernels on the caller's page-locked buffers)
for z in "16 0" "100000 1"; do set -- $z; for w in a1_standard lstm_1x16 a2_full a1_nano; do
NAB200_ZERO_COPY_KFLOATS=$1 NAB200_ASYNC_ZERO_COPY=$2 timeout 200 python bench.py --workload $w --steps 60 --no-cpu-baseline --no-extras --sustained-seconds 0 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('zero-copy=$2 $w dev', round(d['ms_per_step']*1000,1), 'us | async e2e', round(d['e2e']['value']/1e9,3), '| blocking', round(d['e2e']['blocking_value']/1e9,3), 'Gs/s', round(d['e2e']['blocking_ms_per_step']*1000,1),'us parity', d.get('parity',{}).get('max_abs'))"
done; done
