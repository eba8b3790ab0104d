#!/bin/bash
# round-2 fuzz sweep over the rewritten entropy kernels: every entropy mode, fixtures + synthetic shapes + larger files
for mode in 0 1 2; do
  python tools/fuzz_hunt.py --seeds 40 --first 200000 --mode $mode --native 1 2>&1 | tail -1
  python tools/fuzz_hunt.py --synth 1 --seeds 20 --first 210000 --mode $mode 2>&1 | tail -1
  python tools/fuzz_hunt.py --synth 2 --seeds 12 --first 220000 --mode $mode --structural 8 2>&1 | tail -1
done
ls gpurun_out/fuzz 2>/dev/null | head
