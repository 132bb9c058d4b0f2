for t in "" _ng3 _pair _ng3pair; do
  export NFB_LIB_PATH=$PWD/nerfool_b200/libnerfool_b200$t.so
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-nrand --no-bf16 > gpurun_out/exp1$t.json 2> gpurun_out/exp1$t.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/exp1$t.json'))
    print('VARIANT[$t]', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items() if v>1})
except Exception as e:
    print('VARIANT[$t] failed', e); print(open('gpurun_out/exp1$t.err').read()[-800:])
PY
done
