#!/bin/bash
# Run on the GPU box (through gpurun): compute-sanitizer racecheck over the IBRNet parity suite of the final round-2 build.
set -u
out=gpurun_out
mkdir -p $out
SEL="${1:-not full_size and not reference and not properties and not baseline}"
( time timeout 840 compute-sanitizer --tool racecheck --racecheck-report all --launch-timeout 600 --error-exitcode 0 --print-limit 40 \
    python -m pytest tests/test_gpu_parity.py tests/test_aux_losses_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" 2>&1 ) > $out/r02p_racecheck_full.log 2>&1
grep -a "RACECHECK SUMMARY\|passed\|failed\|real" $out/r02p_racecheck_full.log > $out/r02p_racecheck.txt
grep -a -A2 "hazard detected" $out/r02p_racecheck_full.log | grep -a "Write Thread\|Read Thread" | sed 's/.*at //' | sed 's/+0x[0-9a-f]* in / in /' | sort | uniq -c | sort -rn | head -30 >> $out/r02p_racecheck.txt
cat $out/r02p_racecheck.txt
