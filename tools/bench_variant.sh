#!/bin/bash
# usage: tools/bench_variant.sh <lib.so> [...]: short resident-tables bench per kernel-variant build (experiments only)
for lib in "$@"; do
  SC_LIB=$PWD/$lib python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms kernel', round(d['config']['kernel_ms_per_step'],3), d['config']['round_ms'])"
done
