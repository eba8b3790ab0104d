#!/bin/bash
# fuzz sweep after the K2 changes of round 2's second session (sparse-block IDCT, fused kernel on planar-layout frames,
# tile fetch by the last warp, native planes on demand): fixtures + synthetic shapes + larger files, every entropy
# mode, RGBA and native variant, against the oracle
S=${1:-30}
for mode in 0 1 2; do
  python tools/fuzz_hunt.py --seeds $S --first 400000 --mode $mode --native 1 2>&1 | tail -1
  python tools/fuzz_hunt.py --synth 1 --seeds $S --first 410000 --mode $mode --native 1 --structural 4 2>&1 | tail -1
done
python tools/fuzz_hunt.py --synth 2 --seeds $((S/2)) --first 420000 --structural 8 --native 1 2>&1 | tail -1
python tools/fuzz_hunt.py --prog-only 1 --seeds $S --first 430000 2>&1 | tail -1
python tools/fuzz_hunt.py --prog-only 1 --synth 1 --seeds $S --first 440000 --structural 6 2>&1 | tail -1
ls gpurun_out/fuzz 2>/dev/null | head
