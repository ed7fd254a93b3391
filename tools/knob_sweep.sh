#!/bin/bash
for defs in "" "MCIG_SYM_I2F=1" "MCIG_SYM_I2F=1;MCIG_PHILOX_ROUNDS=7"; do
  echo "== $defs"
  for cfg in "65536 128" "65536 512" "75776 128" "303104 256"; do set -- $cfg; MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 20000 $1 $2; done
done
