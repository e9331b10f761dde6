#!/bin/bash
# tools/build_variant.sh NAME [-DKNOB=VALUE ...]  ->  variants/librr_NAME.so (A/B kernel builds; select with bench.py --lib)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart static "$@" \
  -o variants/librr_$name.so brax_rodent_run_b200/csrc/rr_api.cu
echo variants/librr_$name.so
