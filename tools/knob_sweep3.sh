#!/bin/bash
# persistent-CTA size of the dynamically scheduled kernel
for bs in 128 64 32; do
  echo "== MCIG_DYN_BS=$bs"
  MCIG_DYN_BS=$bs python tools/profile_walk.py 100000 65536 0 1
done
echo "== default policy"
for w in 65536 100000 131072; do python tools/profile_walk.py 100000 $w 0 -1; done
MCIG_DYN_BS=32 MCIG_DYN_GRID=2072 python tools/profile_walk.py 100000 65536 0 1
MCIG_DYN_BS=32 MCIG_DYN_GRID=1776 python tools/profile_walk.py 100000 65536 0 1
