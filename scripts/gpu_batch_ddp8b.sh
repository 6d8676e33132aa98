#!/bin/bash
# 8-GPU data-parallel step: one all-reduce after backward vs the windowed schedule (alternating, same box)
N=8; O=gpurun_out
run() { name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp8b_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
l=[x for x in open('gpurun_out/ddp8b_%s.log'%n) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(n, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], flush=True)
else:
    print(n, 'FAILED', open('gpurun_out/ddp8b_%s.log'%n).read()[-600:], flush=True)
PY
}
run window SVIT_DDP_OVERLAP=window
run none SVIT_DDP_OVERLAP=0
run window2 SVIT_DDP_OVERLAP=window
run none2 SVIT_DDP_OVERLAP=0
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp8b_single.log 2>&1; python - <<'PY'
import json
l=[x for x in open('gpurun_out/ddp8b_single.log') if x.startswith('{')]
d=json.loads(l[-1]); print('single', round(d['value']), round(d['ms_per_step'],3))
PY
