#!/bin/bash
# persistent-CTA count of the dynamically scheduled walk kernel at W = 65536 (512 walker blocks of 128 on 148 SMs)
for g in 296 444 480 512; do
  echo "== MCIG_DYN_GRID=$g"
  MCIG_DYN_GRID=$g python tools/profile_walk.py 100000 65536 0 1
done
echo "== static 512"; python tools/profile_walk.py 100000 65536 512 0
echo "== static 128"; python tools/profile_walk.py 100000 65536 128 0
