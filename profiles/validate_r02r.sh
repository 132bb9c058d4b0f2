#!/bin/bash
# Run on the GPU box (through gpurun, ONE GPU): the whole -m gpu suite, the bench lines of every config and the launch lists of the
# final round-2 build (r02r = final build of round 2).  usage: bash profiles/validate_r02r.sh
set -u
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 ) > $out/r02r_gputests.log 2>&1
tail -6 $out/r02r_gputests.log
python bench.py > $out/r02r_bench.json 2> $out/r02r_bench.err; tail -c 400 $out/r02r_bench.json
python bench.py --config 1 --steps 3 --warmup 3 --no-cpu-baseline --no-nrand --no-bf16 > $out/r02r_bench_config1.json 2>> $out/r02r_bench.err
python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-nrand --no-bf16 > $out/r02r_bench_config3.json 2>> $out/r02r_bench.err
python bench.py --config 4 --steps 2 --warmup 3 > $out/r02r_bench_config4_gnt.json 2>> $out/r02r_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02r_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bf16 --no-nrand > $out/r02r_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/r02r_gnt_attack_launches.csv \
    python tests/probes/gnt_attack_probe.py 4096 2 > $out/r02r_gnt_probe.log 2>&1
python - <<'PY'
import json
for f in ('r02r_bench', 'r02r_bench_config1', 'r02r_bench_config3', 'r02r_bench_config4_gnt'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['ms_per_step'], 1), 'ms', round(d['value']), d['unit'], 'e2e', round(d['e2e']['value']), (d.get('attack_step') or {}).get('N_rand_4096', {}).get('ms_per_iter'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
