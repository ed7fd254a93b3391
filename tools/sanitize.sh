#!/bin/bash
# compute-sanitizer memcheck over a cross-section of the GPU tests (every kernel family: register / shared / global walkers,
# dynamic scheduling, replay draws with prefetch padding, estimators incl. the time-split paths, callback, calibration controller)
set -o pipefail
compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 \
  python -m pytest tests -m gpu -x -q -k "c1_simple_short or vec_exp4 or ms_sub16 or ndim_vec256 or ndim_all96 or dep_obs or callback or auto_default or dump or fcblocker_long or mjblocker_large or est_small or est_3d or dynamic_equals_static or device_resident_calibration" 2>&1 | tail -15
