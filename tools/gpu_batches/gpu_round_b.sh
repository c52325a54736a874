#!/bin/sh
# Round B: the GPU tests that failed in round A (after their fixes), then the second column-block sweep (larger windows,
# longer shortest blocked row, and the unblocked plan as the baseline).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/rb_tests.log
grep -E "passed|failed" gpurun_out/rb_tests.log
TAGREC_COLBLOCK=0 python bench.py --steps 3 --no-cpu-baseline --no-c1 --eval-users 0 > gpurun_out/cb_off.json 2> gpurun_out/cb_off.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/cb_off.json").read().strip().split("\n")[-1]); r = d["roofline"]
print(f"colblock off: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  bwd {r['bwd_launch_ms']}", flush=True)
PY
tools/tune_colblock.sh "96 128 160" "256 384 512 768" 2>&1 | tee gpurun_out/colblock_sweep2.txt
