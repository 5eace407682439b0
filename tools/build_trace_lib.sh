#!/bin/sh
# Builds the tracing variant of the engine (per-record and per-phase SM clocks of the first two CTAs, CTA timelines):
#   BLS381_B200_LIB=tools/trace/libbls381_b200_trace.so BLS381_B200_TRACE=<prefix> python bench.py ...
# writes <prefix>_<program>.bin / <prefix>_<program>_ctas.bin; tools/analyze_trace.py summarises them.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/trace
cp -r noble_bls12_381_b200/programs tools/trace/ 2>/dev/null || true
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
    -DBLS381_VM_TRACE -o tools/trace/libbls381_b200_trace.so noble_bls12_381_b200/csrc/api.cu -ldl
