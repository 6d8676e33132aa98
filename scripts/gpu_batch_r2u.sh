#!/bin/bash
# 2-GPU check of the side-stream wgrads under DataParallel: correctness (scripts/check_ddp.py), then the 2-GPU bench with and
# without the side stream.
O=gpurun_out; T=${1:-r2u}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/check_ddp.py 2>&1 | tail -3 | tee $O/${T}_check_ddp.txt
for m in 2 0; do
  SVIT_WGRAD_OVERLAP=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench2_m$m.json 2> $O/${T}_bench2_m$m.err
done
python - $T <<'PY'
import json,sys,glob
T=sys.argv[1]
for f in sorted(glob.glob('gpurun_out/%s_*.json'%T)):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'])
    except Exception as e: print(f, 'failed', e, open(f.replace('.json','.err')).read()[-800:])
PY
