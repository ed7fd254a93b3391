#!/bin/bash
# tail of the walk loop: pre-filter margin with immediates only (MCIG_PREFILTER_SCALED), acceptance counter variants (MCIG_NACC_F64)
run() {
  echo "== defs='$1'"
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 65536 0 1
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 0 1
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 0 0
}
run ""
run "MCIG_PREFILTER_SCALED=1"
run "MCIG_NACC_F64=2"
run "MCIG_PREFILTER_SCALED=1;MCIG_NACC_F64=2"
run "MCIG_PREFILTER_SCALED=1;MCIG_NACC_F64=1"
run ""
