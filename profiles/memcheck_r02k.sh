#!/bin/bash
# Run on the GPU box (through gpurun): compute-sanitizer memcheck over the whole -m gpu suite of the final round-2 build.
set -u
out=gpurun_out
mkdir -p $out
export NFB_SANITIZE=1
( time timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 ) > $out/r02k_memcheck_full.log 2>&1
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned\|real" $out/r02k_memcheck_full.log | tail -12 > $out/r02k_memcheck.txt
cat $out/r02k_memcheck.txt
