#!/bin/bash
# Round-2, last session: one gpurun call on one GPU after the k0_unstuff rewrite -- the GPU test suite, the full bench
# line and the reference arm, a fuzz sweep of the sequential paths (every entropy mode goes through k0), then the launch
# list of the bench command and one full capture of k0_unstuff (each ncu command after the same command exited 0 alone).
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2c_gputests.log; cat $O/r2c_gputests.log
python bench.py > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err
python bench.py --impl reference > $O/r2c_bench_ref_n1.json 2> /dev/null
for mode in 0 1 2; do
  python tools/fuzz_hunt.py --seeds 20 --first 500000 --mode $mode --native 1 2>&1 | tail -1
  python tools/fuzz_hunt.py --synth 1 --seeds 20 --first 510000 --mode $mode --native 1 --structural 4 2>&1 | tail -1
done > $O/r2c_fuzz.log 2>&1
cat $O/r2c_fuzz.log
B="python bench.py --steps 2 --warmup 3 --skip-extras"
$B > $O/r2c_bench_plain.json 2> $O/r2c_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2c_launches_bench.csv $B > $O/r2c_bench_under_ncu.json 2> $O/r2c_bench_under_ncu.err
Q="python tools/quick_bench.py --n 1024 --distinct 64 --iters 1"
$Q > $O/r2c_qb_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k0_unstuff" -c 1 -o $O/r2c_cfg2_k0 -f $Q > $O/r2c_ncu_cfg2_k0.log 2>&1
tail -2 $O/r2c_ncu_cfg2_k0.log; wc -l $O/r2c_launches_bench.csv; head -c 600 $O/r2c_bench_n1.json
