#!/bin/bash
# 8-GPU data-parallel step: overlapped range-wise all-reduce vs one all-reduce after backward, and NCCL CTA caps
O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp8_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
l=[x for x in open('gpurun_out/ddp8_%s.log'%n) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(n, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'])
else:
    print(n, 'FAILED', open('gpurun_out/ddp8_%s.log'%n).read()[-600:])
PY
}
run overlap SVIT_DDP_OVERLAP=1
run nooverlap SVIT_DDP_OVERLAP=0
run overlap_cta4 SVIT_DDP_OVERLAP=1 NCCL_MAX_CTAS=4
run nooverlap_nvls SVIT_DDP_OVERLAP=0 NCCL_ALGO=NVLS
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/ddp8_single.log 2>&1; python - <<'PY'
import json
l=[x for x in open('gpurun_out/ddp8_single.log') if x.startswith('{')]
d=json.loads(l[-1]); print('single', round(d['value']), round(d['ms_per_step'],3))
PY
