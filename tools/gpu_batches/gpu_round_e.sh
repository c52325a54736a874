#!/bin/sh
# Round E: stage isolation of the CTA-pair evaluation kernel (tools/probes/eval_tc2_experiments.py variants).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in 1 2 3 4; do
  echo "== variant t2x$v"
  TAGREC_LIB=$PWD/build/variants/lib_t2x$v.so timeout 200 python tools/eval_bench.py --paths tf32 --reps 5 2>&1 | tail -2 | head -1
done | tee gpurun_out/re_exp.txt
