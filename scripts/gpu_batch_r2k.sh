#!/bin/bash
O=gpurun_out
for i in 1 2 3 4 5 6; do python -m pytest tests/test_gpu_model.py -q -s -k "bookkeeping" 2>&1 | grep -E "un-synchron|passed|failed"; done
python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -2
python scripts/bench_kernels.py gemm 2>&1 | grep -E "gemm_wide|gemm_tn  |dfc1|dqkv|gelu" > $O/r2k_kernels.log; cat $O/r2k_kernels.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2k_bench.log 2>&1
SVIT_NO_WIDE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2k_bench_nowide.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2k_bench2.log 2>&1
python -m pytest tests/test_gpu_model.py tests/test_gpu_callers.py -q -x 2>&1 | tail -2
python - <<'PY'
import json
for f in ('r2k_bench','r2k_bench_nowide','r2k_bench2'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'])
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
