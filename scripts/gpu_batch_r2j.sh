#!/bin/bash
# A/B: polling waits vs hardware-suspended waits (libsvit_b200_hint.so = -DSVIT_WAIT_HINT_NS=100000)
O=gpurun_out
H=$PWD/surface_vision_transformers_b200/libsvit_b200_hint.so
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_bench.log 2>&1
SVIT_LIB=$H python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_bench_hint.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_bench2.log 2>&1
SVIT_LIB=$H python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2j_bench_hint2.log 2>&1
SVIT_LIB=$H python scripts/bench_attn.py > $O/r2j_attn_hint.log 2>&1
python scripts/bench_attn.py > $O/r2j_attn.log 2>&1
SVIT_LIB=$H python scripts/bench_kernels.py gemm > $O/r2j_kernels_hint.log 2>&1
SVIT_LIB=$H python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -2
python - <<'PY'
import json
for f in ('r2j_bench','r2j_bench_hint','r2j_bench2','r2j_bench_hint2'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks'], 'roof', round(d['roofline']['us_per_launch'],1), round(d['roofline_gemm']['us_per_launch'],1))
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
cat $O/r2j_attn.log $O/r2j_attn_hint.log $O/r2j_kernels_hint.log
