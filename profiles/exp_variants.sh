#!/bin/bash
# Record of the build-variant experiment of round 2: the default library against builds with NFB_VTC_NG=3 (3 groups x 168
# registers), NFB_VTC_PAIR=1 (two epilogue chunks per TMEM wait) and both.  The variant libraries are produced with
#   NFB_BUILD_TAG=ng3 NFB_EXTRA_DEFS="-DNFB_VTC_NG=3" python -m nerfool_b200.build      (-> nerfool_b200/libnerfool_b200_ng3.so) ...
# and selected at run time with NFB_LIB_PATH.  Result (ms per step / view forward): default 323.7 / 146.5, ng3 345.6 / 160.2, pair 323.4 / 145.4,
# ng3pair 341.6 / 155.6.  They are not kept in the tree.
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
