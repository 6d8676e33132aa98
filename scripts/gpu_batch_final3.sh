#!/bin/bash
# Final evidence of round 2 after the side-stream / dead-pair / ln_bwd changes: GPU suite, smoke, reference arm, headline bench,
# workload sweep, ncu launch list.  (Section-level ncu of every kernel: profiles/r02d_ncu_all_table.txt, same kernels except
# attn_bwd's dead-pair shortcut and ln_bwd's CTA shape.)
O=gpurun_out; T=${1:-r02i}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -1 | tee $O/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches_sit_small_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu1.log 2>&1
timeout 120 python scripts/bench_attn.py once > $O/${T}_attn_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 1 -o $O/${T}_attn_bwd python scripts/bench_attn.py once > $O/${T}_ncu2.log 2>&1
python scripts/ncu_summary.py $O/${T}_attn_bwd.ncu-rep -o $O/${T}_ncu_attn_bwd_summary.json && python scripts/ncu_hot.py $O/${T}_attn_bwd.ncu-rep 25 > $O/${T}_ncu_attn_bwd.hot.txt 2>/dev/null
rm -f $O/${T}_attn_bwd.ncu-rep
timeout 400 python bench.py --sweep $O/${T}_workloads.json --steps 6 --warmup 3 > $O/${T}_sweep.log 2>&1
python - $T <<'PY'
import json,sys
T=sys.argv[1]
d=json.loads([x for x in open('gpurun_out/%s_bench.json'%T) if x.startswith('{')][-1])
print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'], 'launches', d['gpu_launches'])
print(' roofline', d['roofline']['kernel'][:40], round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), d['roofline'].get('traffic'), round(d['roofline']['step_frac_of_sustained'],4))
print(' cpu', d['cpu_baseline'])
try:
    for r in json.load(open('gpurun_out/%s_workloads.json'%T)):
        print(r['config']['workload'], r['config']['batch_per_gpu'], round(r['value']), round(r['ms_per_step'],2), round(r['roofline']['frac'],3))
except Exception as e: print('sweep', e)
PY
