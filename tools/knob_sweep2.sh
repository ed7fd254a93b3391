#!/bin/bash
# ALU-pipe relief knobs of the headline walk kernel (device/mcig_device.cuh): both default to 1
for defs in "" "MCIG_SYM_MAGIC=0" "MCIG_ACCEPT_FMA=0" "MCIG_SYM_MAGIC=0;MCIG_ACCEPT_FMA=0"; do
  echo "== $defs"
  for cfg in "65536 512" "65536 128" "303104 256"; do set -- $cfg; MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 100000 $1 $2; done
done
