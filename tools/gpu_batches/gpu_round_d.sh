#!/bin/sh
# Round D: the CTA-pair evaluation kernel — parity tests first, then A/B timing against the single-CTA kernel, then the
# cycle accounting of the instrumented variant (tools/probes/eval_tc2_experiments.py 5) if it was built.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "k3_tensor_core" 2>&1 | grep -v Warning | tail -25 > gpurun_out/rd_tests.log
tail -3 gpurun_out/rd_tests.log
if grep -q "failed\|error\|Error" gpurun_out/rd_tests.log; then exit 1; fi
for m in 0 auto; do
  echo "== TAGREC_EVAL_CG2=$m"
  TAGREC_EVAL_CG2=$m timeout 300 python tools/eval_bench.py --paths tf32,fp32 --reps 5 2>&1 | tail -4 | cut -c1-150
done | tee gpurun_out/rd_ab.txt
if [ -f build/variants/lib_t2x5.so ]; then
  TAGREC_LIB=$PWD/build/variants/lib_t2x5.so timeout 200 python tools/eval_bench.py --paths tf32 --reps 1 2>&1 | grep -E "^t2 block 0" | sort | tee gpurun_out/rd_prof.txt
fi
