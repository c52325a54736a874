#!/bin/sh
# Round Q (1 GPU): last forward layer on the batch rows only — full GPU suite, then the bench with and without it.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/rq_tests.log
grep -E "passed|failed|error" gpurun_out/rq_tests.log
for v in 1 0; do
  TAGREC_LAST_LAYER_ROWS=$v python bench.py --steps 5 --no-cpu-baseline --no-c1 --eval-users 0 > gpurun_out/rq_rows$v.json 2> gpurun_out/rq_rows$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/rq_rows{v}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"last_layer_rows={v}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f} x{r['launches_timed']}  rows-launch {r.get('fwd_last_layer_rows_ms')}  bwd {r['bwd_launch_ms']}  loss {d['check']['last_loss']} {d['check']['param_abs_sum']}", flush=True)
except Exception as e:
    print(f"last_layer_rows={v}: FAILED {e}", flush=True); print(open(f"gpurun_out/rq_rows{v}.err").read()[-2000:])
PY
done
