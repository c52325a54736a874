#!/bin/sh
# Round P (1 GPU): the push form of the first backward launch — full GPU suite, then the bench with and without it.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/rp_tests.log
grep -E "passed|failed|error" gpurun_out/rp_tests.log
for pb in 1 0; do
  TAGREC_PUSH_BWD=$pb python bench.py --steps 5 --no-cpu-baseline --no-c1 --eval-users 0 > gpurun_out/rp_push$pb.json 2> gpurun_out/rp_push$pb.err
  python - $pb <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/rp_push{v}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"push={v}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  bwd {r['bwd_launch_ms']}  loss {d['check']['last_loss']} {d['check']['param_abs_sum']}", flush=True)
except Exception as e:
    print(f"push={v}: FAILED {e}", flush=True); print(open(f"gpurun_out/rp_push{v}.err").read()[-1500:])
PY
done
