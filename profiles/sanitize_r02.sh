#!/bin/bash
# Run on the GPU box (through gpurun): compute-sanitizer memcheck and racecheck over the GPU test-suite.
# Keeps the tool's own summaries (tracked copies go to profiles/r02_memcheck.txt / r02_racecheck.txt).
# usage: bash profiles/sanitize_r02.sh [pytest -k expression for racecheck]
set -u
out=gpurun_out
mkdir -p $out
SEL="${1:-not full_size and not reference and not properties}"
export NFB_SANITIZE=1
( time timeout 2400 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 ) > $out/r02_memcheck_full.log 2>&1
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned\|real" $out/r02_memcheck_full.log | tail -12 > $out/r02_memcheck.txt
( time timeout 2400 compute-sanitizer --tool racecheck --racecheck-report all --launch-timeout 600 --error-exitcode 0 --print-limit 30 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$SEL" 2>&1 ) > $out/r02_racecheck_full.log 2>&1
grep -a "RACECHECK SUMMARY\|passed\|failed\|hazard\|Race reported\|real" $out/r02_racecheck_full.log | sort | uniq -c | sort -rn | head -20 > $out/r02_racecheck.txt
tail -3 $out/r02_memcheck.txt; tail -5 $out/r02_racecheck.txt
