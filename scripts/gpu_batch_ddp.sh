#!/bin/bash
# N-GPU data-parallel step: all-reduce schedules (none / window / range), NCCL stream priority.  usage: gpu_batch_ddp.sh N
N=${1:-2}; O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp${N}_$name.log 2>&1
  python - "$name" $N <<'PY'
import json,sys
n=sys.argv[1]; N=sys.argv[2]
l=[x for x in open('gpurun_out/ddp%s_%s.log'%(N,n)) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(n, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], flush=True)
else:
    print(n, 'FAILED', open('gpurun_out/ddp%s_%s.log'%(N,n)).read()[-600:], flush=True)
PY
}
run none SVIT_DDP_OVERLAP=0
run window SVIT_DDP_OVERLAP=window
run window_lowprio SVIT_DDP_OVERLAP=window SVIT_NCCL_HIGH_PRIO=0
run range SVIT_DDP_OVERLAP=1
run none2 SVIT_DDP_OVERLAP=0
run window2 SVIT_DDP_OVERLAP=window
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp${N}_single.log 2>&1; python - $N <<'PY'
import json,sys
l=[x for x in open('gpurun_out/ddp%s_single.log'%sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1]); print('single', round(d['value']), round(d['ms_per_step'],3))
PY
