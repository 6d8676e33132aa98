#!/bin/bash
# Weight-gradient GEMMs on a side stream under the LayerNorm backward kernels: parity (full GPU suite in mode 2, model tests in
# mode 1), then A/B of SVIT_WGRAD_OVERLAP = 0 / 1 / 2 on the headline workload (alternating) and on two other workloads.
O=gpurun_out; T=${1:-r2s}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/${T}_tests.log
SVIT_WGRAD_OVERLAP=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2 | tee $O/${T}_tests_mode1.log
for i in 1 2; do for m in 0 1 2; do
  SVIT_WGRAD_OVERLAP=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_m${m}_$i.json 2> $O/${T}_bench_m${m}_$i.err
done; done
for wl in sit_tiny_ico2_scan_age_train sit_small_ico2_mpp_pretrain; do for m in 0 2; do
  SVIT_WGRAD_OVERLAP=$m timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_${wl}_m${m}.json 2> $O/${T}_${wl}_m${m}.err
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
