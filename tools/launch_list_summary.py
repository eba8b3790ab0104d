"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and average duration per
(kernel, grid), and the share of each kernel inside the device-resident step.

  python tools/launch_list_summary.py gpurun_out/final_launches.csv profiles/r1_launches_bench_summary.txt
"""
import csv
import sys
from collections import OrderedDict


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(src)) if r and r[0] == "ID")
    ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        ms = float(r[iv]) / 1e6
        name = r[ik]
        if name.startswith("void k2_fused"):  # same persistent grid for every batch size: tell the step from the chunks
            name += "  [1024-image step]" if ms > 2.0 else "  [e2e chunks]"
        agg.setdefault((name, r[ig]), []).append(ms)
    lines = [f"# launch list of {src.split('/')[-1]} (ncu --metrics gpu__time_duration.sum --clock-control none; cold caches,",
             "# serialised launches: compare the kernels' SHARES of a step with the CUDA-event timing, not the absolutes)",
             f"{'kernel':58s} {'grid':>14s} {'launches':>8s} {'avg ms':>9s}"]
    for (k, g), v in agg.items():
        lines.append(f"{k[:58]:58s} {g:>14s} {len(v):8d} {sum(v) / len(v):9.3f}")
    # kernels of the device-resident step: the unstuffing pass and the lane kernel on the full batch's grid, the fused kernel
    big = {}
    for (k, g), v in agg.items():
        avg = sum(v) / len(v)
        if k.endswith("[1024-image step]") or ((k.startswith("void k1_lane") or k.startswith("k0_unstuff")) and avg > (1.0 if "k1_lane" in k else 0.2)):
            big[k] = avg
    tot = sum(big.values())
    if tot:
        lines.append("")
        lines.append("device-resident step (1024 images) under ncu: " + ", ".join(f"{k.split('(')[0].replace('void ', '')} {v:.3f} ms = {100 * v / tot:.1f}%" for k, v in big.items()))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
