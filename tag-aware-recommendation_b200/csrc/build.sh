#!/bin/sh
# Build libtagrec_b200.so for sm_100a, in-tree (travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --threads 0 \
    -Xcompiler -fPIC -shared -Xptxas -v \
    api.cu spmm.cu bpr.cu csr_build.cu eval_topk.cu eval_tc.cu eval_tc2.cu eval_auc.cu eval_auc_tc.cu ngcf_dense.cu routing.cu nbr_attention.cu tgcn_tail.cu tgcn_tail_tc.cu tgcn_mix.cu xty.cu sampler.cu adam.cu \
    -o ../libtagrec_b200.so "$@"
