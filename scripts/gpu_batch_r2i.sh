#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -x -k "attn or attention or gemm_ln" 2>&1 | tail -2 > $O/r2i_tests.log
python scripts/bench_attn.py > $O/r2i_attn.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2i_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02b_launches_sit_small_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2i_ncu1.log 2>&1
python scripts/ncu_top.py > $O/r2i_top_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r02_top python scripts/ncu_top.py > $O/r2i_ncu2.log 2>&1
python scripts/ncu_summary.py $O/r02_top.ncu-rep -o $O/r02_top_summary.json; python scripts/ncu_table.py $O/r02_top_summary.json
cat $O/r2i_tests.log $O/r2i_attn.log; ls -la $O/*.ncu-rep $O/*.csv
