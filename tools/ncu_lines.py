#!/usr/bin/env python3
"""Per-SOURCE-LINE view of an ncu report: joins the SASS rows of `ncu --page source --csv` (samples, executed
instructions, active threads) with the line table of the shipped library (`nvdisasm -g`), here on the CPU box.

usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [--kernel REGEX] [--index N] [--top N] [--lib path/to/lib.so]

The library must be the build that was profiled (the SASS rows are matched to the disassembly by position and opcode).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def arg(name, default=None):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def disassemble(lib):
    tmp = tempfile.mkdtemp(prefix="ncu_lines_")
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    best = max((os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")), key=os.path.getsize)
    return subprocess.run(["nvdisasm", "-g", "-c", best], capture_output=True, text=True).stdout


def kernel_lines(dis, mangled_re):
    """[(opcode, file, line)] of the first .text section whose name matches"""
    out, on, cur = [], False, ("?", 0)
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            if on:
                break
            on = re.search(mangled_re, ln) is not None
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            ins = m.group(1).strip()
            parts = ins.split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            out.append((op, cur[0], cur[1]))
    return out


def main():
    rep = sys.argv[1]
    top = int(arg("--top", "40"))
    lib = arg("--lib", os.path.join(ROOT, "tokenizer-zig_b200", "_lib", "libtokzig_b200.so"))
    kre = arg("--kernel", "onepass_kernel")
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"]
    rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
    # several captured launches match: --index N picks the N-th (0-based) of them
    starts = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        sys.exit("no captured kernel matches --kernel")
    rows = rows[starts[min(int(arg("--index", "0")), len(starts) - 1)]:]
    name = rows[0][1]
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    sass = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        try:
            parts = r[ci["Source"]].split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            sass.append((op, int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), float(r[ci["Avg. Threads Executed"]] or 0)))
        except (ValueError, IndexError):
            pass
    # template arguments -> mangled fragment: match on the plain function name and pick the section with the same length
    base = re.split(r"[<(]", name)[0].split("::")[-1].split()[-1]
    dis = disassemble(lib)
    cands = []
    for m in re.finditer(r"//-+ \.text\.(\S*%s\S*) -+" % re.escape(base), dis):
        kl = kernel_lines(dis, re.escape(m.group(1)))
        cands.append((abs(len(kl) - len(sass)), sum(1 for a, b in zip(kl, sass) if a[0] != b[0]), m.group(1), kl))
    cands.sort(key=lambda c: (c[0], c[1]))
    if not cands:
        sys.exit("kernel not found in the library")
    _, mism, mangled, kl = cands[0]
    print(f"## {name}\n## matched {mangled}: {len(sass)} SASS rows, {len(kl)} disassembled, {mism} opcode mismatches")
    n = min(len(kl), len(sass))
    ti = sum(s[2] for s in sass) or 1
    ts = sum(s[1] for s in sass) or 1
    agg = collections.defaultdict(lambda: [0, 0, 0.0])
    for k in range(n):
        key = (kl[k][1], kl[k][2])
        a = agg[key]
        a[0] += sass[k][2]; a[1] += sass[k][1]; a[2] += sass[k][3] * sass[k][2]
    src_cache = {}

    def text(f, l):
        if f not in src_cache:
            p = None
            for d in ("tokenizer-zig_b200/csrc", "tokenizer-zig_b200/host"):
                q = os.path.join(ROOT, d, f)
                if os.path.exists(q):
                    p = q
            src_cache[f] = open(p).read().splitlines() if p else []
        s = src_cache[f]
        return s[l - 1].strip()[:110] if 0 < l <= len(s) else ""

    print(f"## total: {ti} warp instructions, {ts} samples")
    print("## by source line, sorted by executed warp instructions")
    for (f, l), (i, s, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{i / ti * 100:5.1f}% inst {s / ts * 100:5.1f}% smp thr={thr / max(i, 1):4.1f}  {f}:{l}: {text(f, l)}")
    print("## by source line, sorted by stall samples")
    for (f, l), (i, s, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
        print(f"{s / ts * 100:5.1f}% smp {i / ti * 100:5.1f}% inst  {f}:{l}: {text(f, l)}")


if __name__ == "__main__":
    main()
