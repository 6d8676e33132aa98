#!/bin/bash
# A/B of the dead-pair shortcut in the attention backward (SVIT_ATTN_DEBUG=8 = shortcut off), kernel parity tests, short bench.
O=gpurun_out; T=${1:-r2r}
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attn or attention" 2>&1 | tail -3 | tee $O/${T}_tests.log
for i in 1 2; do
  SVIT_ATTN_DEBUG=8 timeout 120 python scripts/bench_attn.py 2>&1 | sed 's/^/off  /' | tee -a $O/${T}_attn.log
  timeout 120 python scripts/bench_attn.py 2>&1 | sed 's/^/on   /' | tee -a $O/${T}_attn.log
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
SVIT_ATTN_DEBUG=8 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_off.json 2> $O/${T}_bench_off.err
python - $T <<'PY'
import json,sys
T=sys.argv[1]
for n in ('bench','bench_off'):
    try:
        d=json.loads([x for x in open('gpurun_out/%s_%s.json'%(T,n)) if x.startswith('{')][-1])
        print(n, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'], round(d['roofline']['us_per_launch'],1))
    except Exception as e: print(n, 'failed', e)
PY
tail -3 $O/${T}_bench.err
