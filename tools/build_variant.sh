#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags for tail.cu ...]: builds build/variants/lib<name>.so (experiments only)
set -e
name=$1; shift
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
compact="-DFR_COMPACT"
for a in "$@"; do [ "$a" = "-DFR_COMPACT_OFF" ] && compact=""; done
nvcc $F $compact "$@" -c -o build/variants/tail_$name.o sumcheck_b200/csrc/tail.cu
nvcc $F -shared -o build/variants/lib$name.so build/sumcheck.o build/variants/tail_$name.o
echo built build/variants/lib$name.so
