#!/bin/bash
# one GPU call: full GPU test suite, micro-benchmarks, step bench, attention-backward timeline, ncu of the attention forward
O=gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v Warning | tail -25 > $O/r2e_tests.log
python scripts/bench_attn.py > $O/r2e_attn.log 2>&1
SVIT_ATTN_FWD_V1=1 python scripts/bench_attn.py > $O/r2e_attn_v1.log 2>&1
python scripts/bench_kernels.py gemm > $O/r2e_kernels.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2e_bench.log 2>&1
SVIT_NO_FUSE_LN=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2e_bench_nofuse.log 2>&1
python scripts/prof_attn_bwd.py > $O/r2e_bwd_timeline.log 2>&1
python scripts/bench_attn.py once > $O/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -o $O/r2e_attn_fwd python scripts/bench_attn.py once > $O/r2e_ncu.log 2>&1
tail -6 $O/r2e_tests.log; cat $O/r2e_attn.log $O/r2e_attn_v1.log $O/r2e_kernels.log; python - <<'PY'
import json
for f in ('gpurun_out/r2e_bench.log', 'gpurun_out/r2e_bench_nofuse.log'):
    l=[x for x in open(f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, 'bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    else:
        print(f, open(f).read()[-1500:])
PY
cat $O/r2e_bwd_timeline.log | head -30
