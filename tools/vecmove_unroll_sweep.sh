#!/bin/bash
# single-index moves beyond ndim 64: unroll cap of the O(ndim) loops (block-end settle of the lazy accumulator, init / finish)
for defs in "" "MCIG_UNROLL_MAX=256" "MCIG_UNROLL_MAX=16"; do
  echo "== defs='$defs'"
  MCIG_JIT_DEFINES="$defs" python -c "
import sys; sys.path.insert(0, 'tools'); import bench_suite as b
b.c3_ndim('vec', ndims=(64, 96, 128, 192, 256), nmc=4000)
" | cut -c1-130
done
