#!/bin/bash
# same-box A/B of library builds over bench workloads: tools/wl_ab.sh "<libs>" "<workloads>"
for r in 1 2; do for lib in $1; do for w in $2; do
NAB200_LIBNAME=$lib timeout 300 python bench.py --no-cpu-baseline --workload $w --steps 60 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('$lib $w', round(d['ms_per_step']*1000,1), 'us', round(d['value']/1e9,3), 'Gs/s', d['clocks']['reasons'])"
done; done; done
