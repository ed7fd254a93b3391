#!/bin/bash
for defs in "" "MCIG_PHILOX_ROUNDS=7" "MCIG_EXP_ESTRIN=1" "MCIG_PHILOX_ROUNDS=7;MCIG_EXP_ESTRIN=1" "MCIG_PHILOX_ROUNDS=2" "MCIG_PHILOX_ROUNDS=2;MCIG_EXP_ESTRIN=1"; do
  echo "== $defs"
  for cfg in "65536 128" "65536 512" "75776 128" "303104 128"; do set -- $cfg; MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 20000 $1 $2; done
done
