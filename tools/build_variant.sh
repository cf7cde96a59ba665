#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags ...]: builds build/variants/lib<name>.so with the extra flags applied
# to all translation units (experiments only; bench with tools/bench_variant.sh or SC_LIB=<path>)
set -e
name=$1; shift
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
compact="-DFR_COMPACT -DFR_INLINE_WIDE_MAC"
for a in "$@"; do [ "$a" = "-DFR_COMPACT_OFF" ] && compact=""; done
nvcc $F "$@" -c -o build/variants/sumcheck_$name.o sumcheck_b200/csrc/sumcheck.cu &
nvcc $F $compact "$@" -c -o build/variants/tail_$name.o sumcheck_b200/csrc/tail.cu &
nvcc $F "$@" -c -o build/variants/gemm_$name.o sumcheck_b200/csrc/gemm.cu &
wait
nvcc $F -shared -o build/variants/lib$name.so build/variants/sumcheck_$name.o build/variants/tail_$name.o build/variants/gemm_$name.o
rm -f build/variants/sumcheck_$name.o build/variants/tail_$name.o build/variants/gemm_$name.o
echo built build/variants/lib$name.so
