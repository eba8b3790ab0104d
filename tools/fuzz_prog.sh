#!/bin/bash
# fuzz sweep over the progressive kernels (lane per scan: zpx_k3l.cu; warp per scan: zpx_k3.cu): progressive fixtures and
# synthetic progressive files, entropy-coded damage + marker-level damage, compared with the oracle
for pm in 0 1; do
  python tools/fuzz_hunt.py --prog-only 1 --prog-mode $pm --seeds ${1:-60} --first 300000 --native 1 2>&1 | tail -1
  python tools/fuzz_hunt.py --prog-only 1 --prog-mode $pm --synth 1 --seeds ${1:-60} --first 310000 --structural 6 2>&1 | tail -1
  python tools/fuzz_hunt.py --prog-only 1 --prog-mode $pm --synth 2 --seeds 30 --first 320000 --structural 8 2>&1 | tail -1
done
ls gpurun_out/fuzz 2>/dev/null | head
