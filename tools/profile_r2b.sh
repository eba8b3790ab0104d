#!/bin/bash
# Round-2 (second session) profiling run: one gpurun call, one GPU.  Every ncu command runs only after the same
# command line has exited 0 without ncu.  Outputs go to gpurun_out/; summaries are made afterwards with
# tools/ncu_summary.py and tools/launch_list_summary.py.
set -u
O=gpurun_out
python bench.py > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err          # the full line (value, e2e, parity, other_configs ...)
python bench.py --impl reference > $O/r2b_bench_ref_n1.json 2> /dev/null  # the reference arm on the same box
B="python bench.py --steps 2 --warmup 3 --skip-extras"
$B > $O/r2b_bench_plain.json 2> $O/r2b_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2b_launches_bench.csv $B > $O/r2b_bench_under_ncu.json 2> $O/r2b_bench_under_ncu.err
Q="python tools/quick_bench.py --n 1024 --distinct 64 --iters 1"
$Q > $O/r2b_qb_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_fused" -c 1 -o $O/r2b_cfg2_k2 -f $Q > $O/r2b_ncu_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_lane" -c 1 -o $O/r2b_cfg2_k1 -f $Q > $O/r2b_ncu_cfg2_k1.log 2>&1
Q4="python tools/quick_bench.py --n 128 --distinct 16 --w 3840 --h 2160 --sub 4:2:2 --iters 1"
$Q4 > $O/r2b_qb_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_fused" -c 1 -o $O/r2b_cfg4_k2 -f $Q4 > $O/r2b_ncu_cfg4.log 2>&1
Q3="python tools/config_bench.py --only cfg3 --iters 1"
$Q3 > $O/r2b_cb_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k2_fused" -c 2 -o $O/r2b_cfg3_k2 -f $Q3 > $O/r2b_ncu_cfg3.log 2>&1
P="python tools/prog_bench.py --n 2048 --skip-warp"
$P > $O/r2b_prog.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_prog.csv $P > $O/r2b_prog_under_ncu.log 2>&1
P2="python tools/prog_bench.py --n 256 --skip-warp"
ncu --set full --clock-control none --import-source on -k regex:"k3l_" -c 9 -o $O/r2b_prog_k3l -f $P2 > $O/r2b_ncu_prog.log 2>&1
tail -2 $O/r2b_ncu_cfg2.log $O/r2b_ncu_cfg4.log $O/r2b_ncu_cfg3.log $O/r2b_ncu_prog.log
wc -l $O/r2b_launches_bench.csv $O/r2b_launches_prog.csv
