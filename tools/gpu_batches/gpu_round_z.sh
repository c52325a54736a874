#!/bin/sh
# Round Z: ncu launch list of the FINAL step structure on lightgcn_1b (2 steps of timed region A).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c1 --eval-users 0"
$CMD > gpurun_out/rz_plain.json 2> gpurun_out/rz_plain.err &&
TAGREC_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_1b_final.csv $CMD > gpurun_out/rz_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_1b_final.csv 2 16
python - <<'PY'
import json
d = json.loads(open("gpurun_out/rz_plain.json").read().strip().split("\n")[-1])
print("plain run:", round(d["ms_per_step"], 1), "ms/step")
PY
