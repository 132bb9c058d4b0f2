#!/bin/bash
# Run on the GPU box (through gpurun): one `ncu --set full` capture of the four tensor-core kernels of a PGD step
# (16384-ray chunks), exported to CSV next to a launch list of a whole default-size step.
# usage: bash profiles/capture.sh <tag>
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:'k_view_tc|k_ray_tc' -s 8 -c 8 -f -o /tmp/prof \
    python bench.py --steps 1 --warmup 3 --max-rays 16384 --no-cpu-baseline --no-bf16 --no-nrand > $out/${tag}_ncu_bench.log 2>&1
ncu -i /tmp/prof.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page details --csv > $out/${tag}_ncu_full_details.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page source --csv --print-source sass > $out/${tag}_ncu_source_sass.csv 2>/dev/null
ls -la /tmp/prof.ncu-rep $out | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bf16 --no-nrand > $out/${tag}_launches_bench.log 2>&1
tail -c 600 $out/${tag}_launches_bench.log
