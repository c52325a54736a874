#!/bin/sh
# Round C: c2 TGCN test detail; then the shipped K1 configuration on lightgcn_1b: plain run, ncu launch list, ncu --set full.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2 3; do
python -m pytest tests/test_gpu_shapes.py -m gpu -q -k c2_tgcn 2>&1 | grep -E "^E  |passed|failed" | cut -c1-400 | head -8
done > gpurun_out/rc_c2.log 2>&1
cat gpurun_out/rc_c2.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c1 --eval-users 0"
$CMD > gpurun_out/rc_plain.json 2> gpurun_out/rc_plain.err &&
TAGREC_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_1b.csv $CMD > gpurun_out/rc_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_1b.csv 2 14
TAGREC_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmm_kernel -c 6 \
    -o gpurun_out/r2_spmm_1b $CMD > gpurun_out/rc_ncu2.log 2>&1
ls -la gpurun_out/r2_spmm_1b.ncu-rep
