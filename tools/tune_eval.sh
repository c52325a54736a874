#!/bin/sh
# K3-TC timing experiments: build variants of libtagrec_b200.so with -DTC_EXPERIMENT=<n> / -DTC_TS=<0|1> and time them
# with tools/eval_bench.py (results of experiments != 0 are WRONG on purpose; they isolate one pipeline stage each).
#   1  no candidate (slow-path) work after the first 8 tiles      -> cost of the exact re-score / K-list path
#   2  drain only half of each accumulator                         -> cost of tcgen05.ld + compare
#   3  no drain at all                                             -> MMA + TMA pipeline floor
#   4  no drain, no TMA                                            -> pure tcgen05.mma issue/execute rate (N = 128)
#   5  as 4 with half as many MMA instructions of N = 256          -> per-instruction overhead of tcgen05.mma
# The experiment branches are NOT in the product sources: they live in tools/probes/eval_tc_experiments.patch (round-1
# kernels) and are applied to a scratch copy of csrc/ by this script.
# usage (on a GPU box): tools/tune_eval.sh 1 3 4 5     (summary of round 1: profiles/r1_eval_tc_experiments.md)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
SRC="api.cu spmm.cu bpr.cu csr_build.cu eval_topk.cu eval_tc.cu eval_tc2.cu eval_auc.cu eval_auc_tc.cu ngcf_dense.cu routing.cu nbr_attention.cu tgcn_tail.cu tgcn_tail_tc.cu tgcn_mix.cu xty.cu sampler.cu adam.cu"
rm -rf build/variants/csrc && cp -r tag-aware-recommendation_b200/csrc build/variants/csrc
mkdir -p build/variants/include && cp include/tagrec_b200.h build/variants/include/
sed -i 's#../../include/tagrec_b200.h#../include/tagrec_b200.h#' build/variants/csrc/common.cuh
(cd build/variants && sed 's#tag-aware-recommendation_b200/##g' ../../tools/probes/eval_tc_experiments.patch | patch -p1 --forward) || \
  { echo "experiment patch no longer applies to the current kernels (it targets the round-1 eval_tc.cu)"; exit 1; }
for n in "$@"; do
  (cd build/variants/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --threads 0 \
     -Xcompiler -fPIC -shared -DTC_EXPERIMENT=$n $SRC -o ../lib_exp$n.so)
  echo "== TC_EXPERIMENT=$n"
  TAGREC_LIB=$PWD/build/variants/lib_exp$n.so python tools/eval_bench.py --paths tf32 --reps 2 2>&1 | tail -3
done
