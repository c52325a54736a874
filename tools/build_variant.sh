#!/bin/sh
# Build a tuning variant of libtagrec_b200.so into build/variants/lib_<name>.so (selected at run time with TAGREC_LIB).
# usage: tools/build_variant.sh l8 -DSPMM_LANES64=8
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
SRC="api.cu spmm.cu bpr.cu csr_build.cu eval_topk.cu eval_tc.cu eval_tc2.cu eval_auc.cu eval_auc_tc.cu ngcf_dense.cu routing.cu nbr_attention.cu tgcn_tail.cu tgcn_tail_tc.cu tgcn_mix.cu xty.cu sampler.cu adam.cu"
(cd tag-aware-recommendation_b200/csrc && ${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
   --threads 0 -Xcompiler -fPIC -shared "$@" $SRC -o ../../build/variants/lib_$name.so)
echo build/variants/lib_$name.so
