#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/r2n_tests.log
python scripts/bench_formats.py 2>&1 | tail -8 | tee $O/r2n_formats.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2n_bench.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2n_ncu1.log 2>&1
python scripts/summarize_launches.py $O/r2n_launches.csv | grep -E "attn_cls|total|gemm_tn_kernel<.*, 1, 0>|gather" 
python - <<'PY'
import json
for f in ('r2n_bench',):
    l=[x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'])
    else:
        print(f, open('gpurun_out/%s.log'%f).read()[-800:])
PY
