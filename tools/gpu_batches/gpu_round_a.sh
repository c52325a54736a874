#!/bin/sh
# One GPU call, several jobs (the pod's queue is the bottleneck): full GPU test suite, the column-block sweep, and the
# 8-lane K1 variants.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/ra_tests.log
grep -E "passed|failed" gpurun_out/ra_tests.log
tools/tune_colblock.sh "32 64 96" "64 128 256" 2>&1 | tee gpurun_out/colblock_sweep.txt
for v in l8 l8u16; do
  for mb in 48; do
    TAGREC_LIB=$PWD/build/variants/lib_$v.so TAGREC_COLBLOCK_MB=$mb python bench.py --steps 3 --no-cpu-baseline --no-c1 --eval-users 0 \
        > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
    python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{v}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"variant {v}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  bwd {r['bwd_launch_ms']}  loss {d['check']['last_loss']}", flush=True)
except Exception as e:
    print(f"variant {v}: FAILED {e}", flush=True)
PY
  done
done 2>&1 | tee gpurun_out/variants.txt
