#!/bin/bash
# PDL A/B + drift probe + full GPU suite
O=gpurun_out
python scripts/probe_drift.py > $O/r2l_drift.log 2>&1; cat $O/r2l_drift.log | tail -40
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/r2l_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_bench_pdl.log 2>&1
SVIT_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_bench_nopdl.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_bench_pdl2.log 2>&1
SVIT_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_bench_nopdl2.log 2>&1
python - <<'PY'
import json
for f in ('r2l_bench_pdl','r2l_bench_nopdl','r2l_bench_pdl2','r2l_bench_nopdl2'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'])
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
