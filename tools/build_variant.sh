#!/bin/bash
# usage: tools/build_variant.sh NAME "<-D flags>" [source.cu ...] : builds csrc/var/lib_NAME.so with the named sources
# (default spx_elementwise.cu) recompiled under the given defines, every other object taken from build/
# (A/B runs: SPX_LIB=.../var/lib_NAME.so python tools/bench_ops.py ...)
set -e
cd "$(dirname "$0")/../shiftedproximaloperators.jl_b200/csrc"
NAME="$1"; DEFS="$2"; shift 2
SRCS="${@:-spx_elementwise.cu}"
mkdir -p var
OBJS=""
for o in build/*.o; do
  b=$(basename "$o" .o)
  if echo " $SRCS " | grep -q " $b.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true \
      -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math $DEFS -c "$b.cu" -o "var/${b}_$NAME.o" 2>/dev/null
    OBJS="$OBJS var/${b}_$NAME.o"
  else
    OBJS="$OBJS $o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "var/lib_$NAME.so" $OBJS -ldl
echo built "var/lib_$NAME.so"
