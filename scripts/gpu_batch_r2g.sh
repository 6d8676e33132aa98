#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -x -k "attn or attention" 2>&1 | tail -3 > $O/r2g_attn_tests.log
python scripts/bench_attn.py > $O/r2g_attn.log 2>&1
SVIT_ATTN_BWD_V1=1 SVIT_ATTN_FWD_V1=1 python scripts/bench_attn.py > $O/r2g_attn_v1.log 2>&1
python scripts/prof_attn_bwd.py > $O/r2g_bwd_timeline.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench.log 2>&1
SVIT_ATTN_BWD_V1=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench_bwdv1.log 2>&1
SVIT_NO_FUSE_LN=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2g_bench_nofuse.log 2>&1
cat $O/r2g_attn_tests.log $O/r2g_attn.log $O/r2g_attn_v1.log
head -16 $O/r2g_bwd_timeline.log
python - <<'PY'
import json
for f in ('r2g_bench','r2g_bench_bwdv1','r2g_bench_nofuse'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'], 'roof', round(d['roofline']['us_per_launch'],1))
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
