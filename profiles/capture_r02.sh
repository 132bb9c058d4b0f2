#!/bin/bash
# Run on the GPU box (through gpurun, ONE GPU): `ncu --set full` capture of the four tensor-core kernels of a PGD step of the
# default bench workload (378x504, 10 source views; 16384-ray chunks), exported to CSV, per-source-line instruction / stall
# tables (profiles/sass_lines.py) and a launch list of a whole default-size step.
# usage: bash profiles/capture_r02.sh <tag> [extra bench args]
set -u
tag=${1:-r02x}
shift
out=gpurun_out
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:'k_view_tc|k_ray_tc' -s 8 -c 8 -f -o /tmp/prof \
    python bench.py --steps 1 --warmup 3 --max-rays 16384 --no-cpu-baseline --no-bf16 --no-nrand "$@" > $out/${tag}_ncu_bench.log 2>&1
ncu -i /tmp/prof.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page details --csv > $out/${tag}_ncu_full_details.csv 2>/dev/null
ncu -i /tmp/prof.ncu-rep --page source --csv --print-source sass > /tmp/${tag}_ncu_source_sass.csv 2>/dev/null
grep -a "^\"Kernel Name\"" /tmp/${tag}_ncu_source_sass.csv | cut -c1-160 > $out/${tag}_source_kernels.txt
B=nerfool_b200/csrc/build
python profiles/sass_lines.py /tmp/${tag}_ncu_source_sass.csv $B/nfb_view_tc_bwd2_inst1.o k_view_tc_fwdILi3ELb1ELb1 1 70 'k_view_tc_fwd' > $out/${tag}_lines_view_fwd.txt 2>&1
python profiles/sass_lines.py /tmp/${tag}_ncu_source_sass.csv $B/nfb_view_tc_bwd2_inst3.o k_view_tc_bwd_stashILi3 1 70 'k_view_tc_bwd_stash' > $out/${tag}_lines_view_bwd.txt 2>&1
python profiles/sass_lines.py /tmp/${tag}_ncu_source_sass.csv $B/nfb_ray_tc_inst5.o k_ray_tcILi3ELb0ELb1 1 50 'k_ray_tc' > $out/${tag}_lines_ray_fwd.txt 2>&1
python profiles/sass_lines.py /tmp/${tag}_ncu_source_sass.csv $B/nfb_ray_tc_inst7.o k_ray_tc_bwd_stashILi3 1 50 'k_ray_tc_bwd_stash' > $out/${tag}_lines_ray_bwd.txt 2>&1
ls -la /tmp/prof.ncu-rep | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bf16 --no-nrand "$@" > $out/${tag}_launches_bench.log 2>&1
tail -c 300 $out/${tag}_launches_bench.log
