#!/usr/bin/env bash
# Developer builds that are NOT part of the product (outputs in benchmarks/_trace/, git-ignored, shipped to the GPU box by gpurun):
#   libsvlm_trace.so  the library with -DSVLM_TRACE (clock64 stamps in the attention kernels; benchmarks/trace_*.py)
#   libsvlm_tune.so   the library with -DSVLM_TUNE  (GEMV configuration chosen from the environment; benchmarks/gemv_tune.py)
#   hw_probe          issue-rate floors of the tcgen05 building blocks (benchmarks/hw_probe.cu)
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p benchmarks/_trace
SRC="abi.cu kv_ops.cu attn_fwd.cu attn_decode.cu gemm.cu gemv.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=false -Xcompiler -fPIC -shared -I../../include"
(cd streaming-vlm_b200/csrc && nvcc $FLAGS -DSVLM_TRACE -o ../../benchmarks/_trace/libsvlm_trace.so $SRC)
(cd streaming-vlm_b200/csrc && nvcc $FLAGS -DSVLM_TUNE -o ../../benchmarks/_trace/libsvlm_tune.so $SRC)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streaming-vlm_b200/csrc -I include benchmarks/hw_probe.cu -o benchmarks/_trace/hw_probe
echo "built benchmarks/_trace/{libsvlm_trace.so,libsvlm_tune.so,hw_probe}"
