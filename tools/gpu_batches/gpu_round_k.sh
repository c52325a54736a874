#!/bin/sh
# Round K (1 GPU): the default bench line exactly as the driver runs it (both arms), then ncu of the CTA-pair evaluation
# kernel (launch list + --set full) on the evaluation micro-benchmark.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/rk_bench_ref.json 2> gpurun_out/rk_bench_ref.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/rk_bench.json 2> gpurun_out/rk_bench.err; echo "our arm rc=$?"
tail -c 600 gpurun_out/rk_bench.err
CMD="python tools/eval_bench.py --paths tf32 --reps 1"
$CMD > gpurun_out/rk_eval_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_tc2_kernel -s 1 -c 1 -o gpurun_out/r2_eval_tc2 $CMD > gpurun_out/rk_ncu.log 2>&1
ls -la gpurun_out/r2_eval_tc2.ncu-rep
