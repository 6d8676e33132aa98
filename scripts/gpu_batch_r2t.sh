#!/bin/bash
# ln_bwd as 128-thread CTAs at 112 registers (three fit next to a wgrad CTA): full GPU suite, then A/B against the former
# 256-thread CTAs (SVIT_LN_BWD_WARPS=8) on the headline workload, alternating, plus the kernel alone (scripts/bench_kernels.py).
O=gpurun_out; T=${1:-r2t}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/${T}_tests.log
for i in 1 2; do for wv in 8 4; do
  SVIT_LN_BWD_WARPS=$wv timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_w${wv}_$i.json 2> $O/${T}_bench_w${wv}_$i.err
done; done
python - $T <<'PY'
import json,sys,glob
T=sys.argv[1]
for f in sorted(glob.glob('gpurun_out/%s_*.json'%T)):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'])
    except Exception as e: print(f, 'failed', e)
PY
