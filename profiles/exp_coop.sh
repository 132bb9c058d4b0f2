#!/bin/bash
# Build-variant experiment: quarter-warp cooperative scatter in the view backward (NFB_VTC_COOP_SCATTER=1) against the default library.
export NFB_LIB_PATH=$PWD/nerfool_b200/libnerfool_b200_cs.so
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200
for t in "" _cs; do
  export NFB_LIB_PATH=$PWD/nerfool_b200/libnerfool_b200$t.so
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-nrand --no-bf16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('VARIANT[$t]', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['kernel_ms_per_step'].items() if v>1}, 'fwd frame', round(d['fwd_ms_per_frame'],1))"
done
