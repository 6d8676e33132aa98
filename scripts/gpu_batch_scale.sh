#!/bin/bash
# 1 / 2 / 4 / 8 GPUs of one box, back to back (bench.py defaults: windowed all-reduce, NCCL on a high-priority stream)
O=gpurun_out; T=${1:-r02c}
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_scale_1gpu.json 2> $O/${T}_scale_1gpu.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_scale_${N}gpu.json 2> $O/${T}_scale_${N}gpu.err
done
python - $T <<'PY'
import json,sys
T=sys.argv[1]; base=None
for N in (1,2,4,8):
    l=[x for x in open('gpurun_out/%s_scale_%dgpu.json'%(T,N)) if x.startswith('{')]
    if not l:
        print(N, 'FAILED', open('gpurun_out/%s_scale_%dgpu.err'%(T,N)).read()[-500:]); continue
    d=json.loads(l[-1]); base = base or d['value']
    print(N, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'eff', round(d['value']/(N*base),4), d['clocks']['sm_mhz'], flush=True)
PY
