#!/bin/bash
# compute-sanitizer memcheck over a cross-section of the GPU tests (every kernel family: register / shared / global walkers,
# dynamic scheduling, replay draws with prefetch padding, estimators incl. the time-split paths, callback, calibration controller)
set -o pipefail
compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 \
  python -m pytest tests -m gpu -x -q -k "c1_simple_short or vec_exp4 or ms_sub16 or ndim_vec256 or ndim_all96 or dep_obs or callback or auto_default or dump or fcblocker_long or mjblocker_large or est_small or est_3d or dynamic_equals_static or device_resident_calibration" 2>&1 | tail -15
# round 2 kernel families: lane-split walkers, MultiStepMove with the committed position out of shared memory, C3 shapes on every placement,
# chunked staging + fold kernels, fused one-pass estimators, lazily accumulated sums, user-defined moves / domains, parameterised proposals
# (closed forms and the fixed-count Marsaglia-Tsang sampler), device-resident control loops
compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 \
  python -m pytest tests -m gpu -x -q -k "c3_shapes_replay or lane_split_walkers_replay or user_defined_domain or chunked_staging_equals or fused_one_pass or lazy_accumulation_against or (user_defined_move and 24) or (parameterised_proposals_follow and gamma) or device_resident_calibration or warp_specialised" 2>&1 | tail -15
