#!/bin/bash
# compute-sanitizer memcheck over a cross-section of the GPU tests: bash tools/sanitize.sh [tag]  (writes gpurun_out/<tag>_sanitizer_{a,b}.log, prints a summary)
# (a) every round-1 kernel family: register / shared / global walkers, dynamic scheduling, replay draws with prefetch padding, estimators incl. the
#     time-split paths, callback, calibration controller
# (b) round 2: lane-split walkers, MultiStepMove with the committed position out of shared memory, C3 shapes on every placement, chunked staging + fold
#     kernels, fused one-pass estimators, lazily accumulated sums, user-defined moves / domains, parameterised proposals (closed forms and the
#     fixed-count Marsaglia-Tsang sampler), device-resident control loops
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
SEL_A="c1_simple_short or vec_exp4 or ms_sub16 or ndim_vec256 or ndim_all96 or dep_obs or callback or auto_default or dump or fcblocker_long or mjblocker_large or est_small or est_3d or dynamic_equals_static or device_resident_calibration"
SEL_B="c3_shapes_replay or lane_split_walkers_replay or user_defined_domain or chunked_staging_equals or fused_one_pass or lazy_accumulation_against or (user_defined_move and 24) or (parameterised_proposals_follow and gamma) or device_resident_calibration or warp_specialised"
for s in a b; do
    sel="$SEL_A"; [ $s = b ] && sel="$SEL_B"
    compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 300 --print-limit 20 \
        python -m pytest tests -m gpu -x -q -k "$sel" > $out/${tag}_sanitizer_$s.log 2>&1
    echo "[$s] exit $?"
    grep -E "passed|failed" $out/${tag}_sanitizer_$s.log | tail -1
    grep -E "^========= (Program hit|Invalid|Misaligned|Error|Leaked|Uninit|Race)" $out/${tag}_sanitizer_$s.log | sed -E 's/0x[0-9a-f]+/0x../g' | sort | uniq -c | sort -rn | head -8
    grep "ERROR SUMMARY" $out/${tag}_sanitizer_$s.log | head -1
    # keep the committed log small: the per-error host backtraces are not evidence
    grep -v "Host Frame" $out/${tag}_sanitizer_$s.log | head -200 > $out/${tag}_sanitizer_$s.txt
done
