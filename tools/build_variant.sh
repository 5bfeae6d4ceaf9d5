#!/bin/bash
# usage: tools/build_variant.sh NAME "<-D flags>" : builds csrc/var/lib_NAME.so with spx_elementwise.cu recompiled
# under the given defines (A/B runs: SPX_LIB=.../var/lib_NAME.so python tools/bench_ops.py ...)
set -e
cd "$(dirname "$0")/../shiftedproximaloperators.jl_b200/csrc"
mkdir -p var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math $2 -c spx_elementwise.cu -o var/ew_$1.o 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o var/lib_$1.so var/ew_$1.o build/spx_context.o build/spx_l1b2.o \
  build/spx_group.o build/spx_topr.o build/spx_host.o
echo built var/lib_$1.so
