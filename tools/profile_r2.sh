#!/bin/bash
# Round-2 profiling run (one gpurun call, one GPU).  Every ncu command runs only after the same command line has
# exited 0 without ncu.  Outputs go to gpurun_out/; summaries are made afterwards with tools/ncu_summary.py.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --skip-extras"
$B > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_bench_under_ncu.json 2> $O/r2_bench_under_ncu.err
Q="python tools/quick_bench.py --n 1024 --distinct 64 --iters 1"
$Q > $O/r2_qb_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k0_unstuff|k1_lane|k2_fused" -c 3 -o $O/r2_cfg2_full -f $Q > $O/r2_ncu_cfg2.log 2>&1
Q4="python tools/quick_bench.py --n 128 --distinct 16 --w 3840 --h 2160 --sub 4:2:2 --iters 1"
$Q4 > $O/r2_qb_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_fused" -c 1 -o $O/r2_cfg4_k2 -f $Q4 > $O/r2_ncu_cfg4.log 2>&1
Q3="python tools/config_bench.py --only cfg3 --iters 1"
$Q3 > $O/r2_cb_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k1s_|k2_fused" -c 7 -o $O/r2_cfg3_full -f $Q3 > $O/r2_ncu_cfg3.log 2>&1
tail -2 $O/r2_ncu_cfg2.log $O/r2_ncu_cfg4.log $O/r2_ncu_cfg3.log
wc -l $O/r2_launches_bench.csv
