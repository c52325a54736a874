#!/bin/sh
# Round W: K1 with the cp.async ring gather (SPMM_ASYNC variants) — parity of the K1 / LightGCN tests, then the bench.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/build/variants/lib_$v.so
  echo "== $v: $(TAGREC_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k 'k1_ or lightgcn or trajectory or ngcf' 2>&1 | tail -1)"
  TAGREC_LIB=$lib timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c1 --eval-users 0 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{v}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"variant {v}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  rows {r.get('fwd_last_layer_rows_ms')}  bwd {r['bwd_launch_ms']}  loss {d['check']['last_loss']} {d['check']['param_abs_sum']}", flush=True)
except Exception as e:
    print(f"variant {v}: FAILED {e}", flush=True); print(open(f"gpurun_out/var_{v}.err").read()[-1500:])
PY
done 2>&1 | tee gpurun_out/variants_w.txt
