#!/usr/bin/env python3
"""Reads an .ncu-rep (ncu --set full) here on the CPU box and prints, per captured kernel: duration, DRAM bytes, key
utilisation / stall metrics, and the SASS instructions that collect the most stall samples.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 8]"""
import collections
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch uniformity %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected")]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 8
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        names.append(name)
        print(f"## {name}")
        for key, label in WANT:
            if key in idx:
                print(f"  {label:32s} {r[idx[key]]:>18s} {units[idx[key]]}")
    for name in dict.fromkeys(names):
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{name.split('::')[-1]}"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        blocks, cur = [], None
        for r in rr:                                     # one block per captured launch; take the first
            if r and r[0] == "Kernel Name":
                cur = []; blocks.append(cur)
            elif cur is not None:
                cur.append(r)
        if not blocks:
            continue
        b = blocks[0]
        h = b[0]; ci = {n: i for i, n in enumerate(h)}
        S, I, SRC = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
        data = []
        for n, r in enumerate(b[1:]):
            try:
                data.append((int(r[S]), int(r[I]), n, r[SRC]))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1
        agg = collections.Counter()
        for s, _i, _n, src in data:
            parts = src.split()
            op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
            agg[op.split(".")[0]] += s
        print(f"## {name}: stall samples by opcode: " + ", ".join(f"{o} {v / tot * 100:.1f}%" for o, v in agg.most_common(8)))
        for s, i, n, src in sorted(data, reverse=True)[:top]:
            print(f"   {s / tot * 100:5.1f}%  sass#{n:<5d} exec={i:<10d} {src[:100]}")


if __name__ == "__main__":
    main()
