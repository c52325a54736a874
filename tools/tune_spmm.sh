#!/bin/sh
# usage: tools_tune.sh <workload> <variants...>  -> one line per variant with the live K1 timings
wl=$1; shift
for v in "$@"; do
  TAGREC_LIB=$PWD/build/variants/lib_$v.so python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --eval-users 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$v', 'step_ms=%.1f fwd_ms=%.2f bwd_ms=%.2f frac=%.3f'%(d['ms_per_step'], r['ms_per_launch'], r['bwd_layer_ms'], r['frac']))
"
done
