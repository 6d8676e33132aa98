#!/bin/bash
# cls-only last block: full GPU suite + A/B bench
O=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/r2m_tests.log
python -m pytest tests/test_gpu_model.py -q -s -k "cls_pooling" 2>&1 | grep -E "cls-only|passed|failed" | tee -a $O/r2m_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2m_bench_cls.log 2>&1
SVIT_FULL_LAST_LAYER=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2m_bench_full.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2m_bench_cls2.log 2>&1
SVIT_FULL_LAST_LAYER=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2m_bench_full2.log 2>&1
python - <<'PY'
import json
for f in ('r2m_bench_cls','r2m_bench_full','r2m_bench_cls2','r2m_bench_full2'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'], round(d['roofline']['step_frac_of_sustained'],4))
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
