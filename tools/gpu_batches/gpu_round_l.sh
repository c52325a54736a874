#!/bin/sh
# Round L: pair-kernel item splits (wave-aware plan) — parity tests, then timing for several split counts.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "k3_tensor_core" 2>&1 | tail -3
for sp in auto 1 2 4 8 15; do
  echo "== TAGREC_EVAL_SPLITS=$sp"
  if [ $sp = auto ]; then unset TAGREC_EVAL_SPLITS; else export TAGREC_EVAL_SPLITS=$sp; fi
  timeout 300 python tools/eval_bench.py --paths tf32,fp32 --reps 5 2>&1 | grep -E "tf32|identical" | cut -c1-150
  timeout 300 python tools/eval_bench.py --users 32768 --paths tf32 --reps 5 2>&1 | grep -E "tf32" | cut -c1-150
done | tee gpurun_out/rl_splits.txt
