#!/bin/bash
# Run on the GPU box (through gpurun): compute-sanitizer memcheck + racecheck over the GNT tests (forward, nfb_gnt_bwd, camera gradients).
set -u
out=gpurun_out
mkdir -p $out
( time timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_gnt_gpu.py -m gpu -q -x -p no:cacheprovider -k "not full_size" 2>&1 ) > $out/r02_gnt_memcheck_full.log 2>&1
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned\|real" $out/r02_gnt_memcheck_full.log | tail -12 > $out/r02_gnt_memcheck.txt
( time timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --launch-timeout 600 --error-exitcode 0 --print-limit 40 \
    python -m pytest tests/test_gnt_gpu.py -m gpu -q -x -p no:cacheprovider -k "backward or camera or attack" 2>&1 ) > $out/r02_gnt_racecheck_full.log 2>&1
grep -a "RACECHECK SUMMARY\|passed\|failed\|real" $out/r02_gnt_racecheck_full.log > $out/r02_gnt_racecheck.txt
grep -a -A2 "hazard detected" $out/r02_gnt_racecheck_full.log | grep -a "Write Thread\|Read Thread" | sed 's/.*at //' | sed 's/+0x[0-9a-f]* in / in /' | sort | uniq -c | sort -rn | head -30 >> $out/r02_gnt_racecheck.txt
cat $out/r02_gnt_memcheck.txt $out/r02_gnt_racecheck.txt
