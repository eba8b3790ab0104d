"""Where the warps of k2_fused spend their time: warp-state samples of an ncu capture (--set full --import-source on)
grouped by the part of the kernel the sampled instruction belongs to (classified by how often it executes per tile)
and by stall reason.
  python tools/k2_stall_breakdown.py gpurun_out/r2b_cfg2_k2.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(rep, out=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = rows[1]
    ia, isamp = h.index("Instructions Executed"), h.index("# Samples")
    reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    data = [r for r in rows[2:] if len(r) > isamp and r[ia].isdigit()]
    tot = sum(int(r[isamp]) for r in data) or 1
    per = defaultdict(lambda: [0, 0, defaultdict(int)])
    for r in data:
        e = per[int(r[ia])]
        e[0] += 1
        e[1] += int(r[isamp])
        for c in reasons:
            v = r[h.index(c)]
            if v.isdigit():
                e[2][c] += int(v)
    lines = [f"# warp-state samples of {rep.split('/')[-1]} by execution count of the sampled instruction ({tot} samples)",
             f"{'executions':>12s} {'instrs':>6s} {'samples':>8s} {'share':>6s}  top stall reasons"]
    for n, (k, sm, rs) in sorted(per.items(), key=lambda kv: -kv[1][1])[:10]:
        top = ", ".join(f"{c[6:]} {100 * v / max(sm, 1):.0f}%" for c, v in sorted(rs.items(), key=lambda kv: -kv[1])[:4])
        lines.append(f"{n:12d} {k:6d} {sm:8d} {100 * sm / tot:5.1f}%  {top}")
    allr = defaultdict(int)
    for _, (_, _, rs) in per.items():
        for c, v in rs.items():
            allr[c] += v
    lines.append("all samples by reason: " + ", ".join(f"{c[6:]} {100 * v / tot:.1f}%" for c, v in sorted(allr.items(), key=lambda kv: -kv[1])[:9]))
    s = "\n".join(lines) + "\n"
    print(s)
    if out:
        open(out, "a").write("\n" + s)


if __name__ == "__main__":
    main(*sys.argv[1:])
