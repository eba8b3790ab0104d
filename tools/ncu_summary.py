"""Summarise an ncu report (.ncu-rep) into a small text file for profiles/ (the judged evidence).

  python tools/ncu_summary.py gpurun_out/prof_k2.ncu-rep profiles/r1_k2_fused.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)"]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        lines.append(f"\n## {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:96s} {row[i]:>16s} {units[i]}")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        data = [r for r in src[2:] if len(r) > isamp and r[ia].isdigit()]
        tot = sum(int(r[ia]) for r in data)
        ts = sum(int(r[isamp]) for r in data) or 1
        ops = Counter()
        for r in data:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";").split(".")[0]
            ops[op] += int(r[ia])
        lines.append(f"\n## source page: {len(data)} SASS instructions, {tot} warp-instructions executed, {ts} stall samples")
        lines.append("opcode mix (executed warp-instructions): " + ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(14)))
        tma = [o for o in ops if o in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS")]
        lines.append("TMA / mbarrier opcodes present: " + (", ".join(f"{o} x{ops[o]}" for o in tma) or "none"))
        lines.append("top stall locations (samples, share, executions, SASS):")
        for r in sorted(data, key=lambda r: -int(r[isamp]))[:12]:
            lines.append(f"  {int(r[isamp]):7d} {100*int(r[isamp])/ts:5.1f}%  {int(r[ia]):9d}  {r[isrc].strip()[:100]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out, len(lines), "lines")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
