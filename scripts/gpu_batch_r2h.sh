#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -3 > $O/r2h_kernel_tests.log
python scripts/bench_attn.py > $O/r2h_attn.log 2>&1
python scripts/prof_attn_bwd.py > $O/r2h_bwd_timeline.log 2>&1
python scripts/bench_kernels.py gemm > $O/r2h_kernels.log 2>&1
SVIT_LN_STAGGER_NS=0 python scripts/bench_kernels.py gemm 2>&1 | grep gemm_ln > $O/r2h_kernels_nostagger.log
SVIT_LN_STAGGER_NS=20000 python scripts/bench_kernels.py gemm 2>&1 | grep gemm_ln > $O/r2h_kernels_stagger20.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2h_bench.log 2>&1
SVIT_LN_STAGGER_NS=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2h_bench_nostagger.log 2>&1
cat $O/r2h_kernel_tests.log $O/r2h_attn.log; head -12 $O/r2h_bwd_timeline.log; cat $O/r2h_kernels.log; echo nostagger; cat $O/r2h_kernels_nostagger.log; echo stagger20us; cat $O/r2h_kernels_stagger20.log
python - <<'PY'
import json
for f in ('r2h_bench','r2h_bench_nostagger'):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'], 'roof', round(d['roofline']['us_per_launch'],1))
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
