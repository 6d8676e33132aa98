#!/bin/bash
# Last evidence batch of the round (the --set full captures of the dominant kernels, r02c, are unchanged since): the GPU suite
# three times in a row (flakiness check), smoke, headline bench + reference arm, workload sweep, launch list, section-level ncu
# of every kernel of the step.
O=gpurun_out; T=${1:-r02d}
for i in 1 2 3; do python -m pytest tests -m gpu -q 2>&1 | tail -1; done | tee $O/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python bench.py --impl reference --steps 10 --warmup 3 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --sweep $O/${T}_workloads.json --steps 10 --warmup 3 > $O/${T}_sweep.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches_sit_small_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu1.log 2>&1
python scripts/ncu_all.py > $O/${T}_all_plain.log 2>&1 && timeout 900 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section WarpStateStats --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sass__inst_executed_local_loads,sass__inst_executed_local_stores --clock-control none --profile-from-start off -o $O/${T}_all python scripts/ncu_all.py > $O/${T}_ncu3.log 2>&1
python scripts/ncu_summary.py $O/${T}_all.ncu-rep -o $O/${T}_ncu_all_summary.json && python scripts/ncu_table.py $O/${T}_ncu_all_summary.json > $O/${T}_ncu_all_table.txt
rm -f $O/${T}_all.ncu-rep
python - $T <<'PY'
import json,sys
T=sys.argv[1]
d=json.loads([x for x in open('gpurun_out/%s_bench.json'%T) if x.startswith('{')][-1])
print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'], 'launches', d['gpu_launches'])
print(' roofline', d['roofline']['kernel'][:40], round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), d['roofline'].get('traffic'), round(d['roofline']['step_frac_of_sustained'],4))
print(' cpu', d['cpu_baseline'])
for r in json.load(open('gpurun_out/%s_workloads.json'%T)):
    print(r['config']['workload'], r['config']['batch_per_gpu'], round(r['value']), round(r['ms_per_step'],2), round(r['roofline']['frac'],3))
PY
grep -E "mpp_loss|cast_bf16|masked_rowsum|gemv|gather|transpose_table|attn_cls|pack_patches" $O/${T}_ncu_all_table.txt | cut -c1-130
