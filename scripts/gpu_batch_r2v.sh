#!/bin/bash
# cls-path wgrads on the side stream: full GPU suite; A/B of the side stream's priority and of programmatic dependent launch.
O=gpurun_out; T=${1:-r2v}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/${T}_tests.log
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_default_$i.json 2> $O/${T}_bench_default_$i.err
  SVIT_SIDE_PRIO=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_lowprio_$i.json 2> $O/${T}_bench_lowprio_$i.err
  SVIT_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_nopdl_$i.json 2> $O/${T}_bench_nopdl_$i.err
done
python - $T <<'PY'
import json,sys,glob
T=sys.argv[1]
for f in sorted(glob.glob('gpurun_out/%s_*.json'%T)):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'])
    except Exception as e: print(f, 'failed', e)
PY
