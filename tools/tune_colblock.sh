#!/bin/sh
# K1 column-blocked plan sweep on lightgcn_1b: window size (MB of 256-byte rows) x shortest blocked row.
# usage (GPU box): tools/tune_colblock.sh "32 48 64 96" "128 256 384"     -> one summary line per combination
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in $1; do for md in $2; do
  TAGREC_COLBLOCK_MB=$mb TAGREC_COLBLOCK_MIN_DEG=$md python bench.py --steps 3 --no-cpu-baseline --no-c1 --eval-users 0 \
      > gpurun_out/cb_${mb}_${md}.json 2> gpurun_out/cb_${mb}_${md}.err
  python - $mb $md <<'PY'
import json, sys
mb, md = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/cb_{mb}_{md}.json").read().strip().split("\n")[-1])
    r = d["roofline"]
    print(f"window {mb} MB min_deg {md}: step {d['ms_per_step']:.1f} ms  fwd {r['ms_per_launch']:.2f}  bwd {r['bwd_launch_ms']}  "
          f"plan {d['run']['plan']}  setup {d['setup_s']}", flush=True)
except Exception as e:
    print(f"window {mb} MB min_deg {md}: FAILED {e}", flush=True)
PY
done; done
