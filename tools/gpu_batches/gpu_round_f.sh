#!/bin/sh
# Round F: cycle accounting variants of the CTA-pair evaluation kernel (tools/probes/eval_tc2_experiments.py 5 6 7).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in "$@"; do
  echo "== variant t2x$v"
  TAGREC_LIB=$PWD/build/variants/lib_t2x$v.so timeout 200 python tools/eval_bench.py --paths tf32 --reps 1 2>&1 | grep -E "^t2 block 0|\"path\"" | sort | awk '/scorer|path/ || NR%3==1' | cut -c1-330
done | tee gpurun_out/rf_prof.txt
