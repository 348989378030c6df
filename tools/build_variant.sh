#!/bin/bash
# tools/build_variant.sh <out.so> [-D...]: an experimental build of the library next to the product one
out="$1"; shift
cd "$(dirname "$0")/../copula-msm-and-copula-garch-var_b200/csrc" && \
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --cudart shared -Xcompiler -fPIC -shared "$@" -o "$out" cvar_api.cu
