#!/bin/bash
# sweep.sh LOG CMD... -- run CMD once per line of knob settings read from stdin ("NAME=VALUE NAME2=VALUE2", empty line = defaults), appending to LOG.
# One parameterised script instead of a knob_sweepN.sh per experiment (the round-1 sweeps, profiles/r01_knob_sweep*.log, were made this way by hand).
log=$1; shift
while IFS= read -r knobs; do
    echo "### ${knobs:-defaults}" | tee -a "$log"
    env $knobs "$@" 2>&1 | tee -a "$log"
done
