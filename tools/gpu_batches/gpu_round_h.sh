#!/bin/sh
# Round H: K1 rolling-gather variants on lightgcn_1b (default plan), against the shipped library in the same call.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in base "$@"; do
  lib=""; [ "$v" != base ] && lib=$PWD/build/variants/lib_$v.so
  TAGREC_LIB=$lib python bench.py --steps 3 --no-cpu-baseline --no-c1 --eval-users 0 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{v}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"variant {v}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  bwd {r['bwd_launch_ms']}  loss {d['check']['last_loss']} {d['check']['param_abs_sum']}", flush=True)
except Exception as e:
    print(f"variant {v}: FAILED {e}", flush=True)
PY
done 2>&1 | tee gpurun_out/variants_h.txt
