#!/bin/bash
# Run on the GPU box (through gpurun): compute-sanitizer racecheck + memcheck over the tests that drive the cooperative gather / scatter
# (fused view forward / backward, standalone projector) of the final round-2 build.
set -u
out=gpurun_out
mkdir -p $out
( time timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --launch-timeout 600 --error-exitcode 0 --print-limit 40 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "projector or render_rays_outputs or golden_end_to_end or stash_backward" 2>&1 ) > $out/r02q_racecheck_full.log 2>&1
grep -a "RACECHECK SUMMARY\|passed\|failed\|real" $out/r02q_racecheck_full.log > $out/r02q_racecheck.txt
grep -a -A2 "hazard detected" $out/r02q_racecheck_full.log | grep -a "Write Thread\|Read Thread" | sed 's/.*at //' | sed 's/+0x[0-9a-f]* in / in /' | sort | uniq -c | sort -rn | head -30 >> $out/r02q_racecheck.txt
( time timeout 300 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 ) > $out/r02q_memcheck_full.log 2>&1
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned\|real" $out/r02q_memcheck_full.log | tail -6 >> $out/r02q_racecheck.txt
cat $out/r02q_racecheck.txt
