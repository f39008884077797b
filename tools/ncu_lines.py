#!/usr/bin/env python
"""Per-kernel, per-source-region view of an `ncu --set full --import-source on` report (kernel built with -lineinfo):
executed warp instructions per SimPy event, static SASS bytes and stall samples of every handler / function region.

usage: python tools/ncu_lines.py REPORT.ncu-rep EVENTS_PER_LAUNCH [--source FILE] [--top N] [--lines LO-HI] [--kernel I]
The kernel is instruction-fetch bound, so the static size of the WARM regions is what a change has to shrink."""
import argparse, collections, csv, re, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("events", type=float)
ap.add_argument("--source", default="/root/repo/manusim_b200/csrc/msim_kernel.cu")
ap.add_argument("--top", type=int, default=40); ap.add_argument("--lines", default=None); ap.add_argument("--kernel", type=int, default=0)
a = ap.parse_args()
out = subprocess.run(f"ncu -i {a.report} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
# a new kernel starts where the main source file's section starts again
main = a.source.split("/")[-1]
kern, k = [], -1
for hi in hdr_idx:
    f = rows[hi - 2][1].split("/")[-1] if hi >= 2 and len(rows[hi - 2]) > 1 else "?"
    if f.endswith(".cu") and ("msim_kernel" in f):
        k += 1
    kern.append((max(k, 0), f, hi))
hdr_idx.append(len(rows) + 2)
srcf = open(a.source).read().split("\n")
marks = [(i + 1, s.strip()[:64]) for i, s in enumerate(srcf) if re.match(r"\s*(case OP_|__device__|template|__global__|default:|L_[A-Z_]+:)", s)]
def region(ln):
    name = "top"
    for ml, ms in marks:
        if ml <= ln: name = f"{ml}:{ms}"
        else: break
    return name
ex, sz, st, per_line = collections.Counter(), collections.Counter(), collections.Counter(), {}
for s, (kk, f, hi) in enumerate(kern):
    if kk != a.kernel: continue
    hdr = rows[hi]; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
    cur = None                                         # the source line the following SASS rows belong to
    for r in rows[hi + 1: hdr_idx[s + 1] - 2]:
        if len(r) <= ie: continue
        if not r[0].isdigit():
            # a SASS row ("", "", address, instruction, ...) under its source line: 16 bytes of code
            if cur is not None and len(r) > 2 and r[2].strip().startswith("0x"):
                sz[cur[0]] += 16
                if cur[1] in per_line: per_line[cur[1]][1] += 16
            continue
        ln = int(r[0]); v = int(r[ie] or 0); smp = int(r[isamp] or 0)
        key = region(ln) if "msim_kernel" in f else f
        cur = (key, ln if "msim_kernel" in f else None)
        ex[key] += v; st[key] += smp
        if "msim_kernel" in f:
            e = per_line.setdefault(ln, [0, 0, 0, r[1]]); e[0] += v; e[2] += smp
T, S, Z = sum(ex.values()), sum(st.values()), sum(sz.values())
print(f"kernel {a.kernel}: {T / a.events:.1f} warp instr/event, {Z / 1024:.1f} KB SASS, {S} stall samples")
print(" instr/ev   KB   stall%  region")
for kx, v in ex.most_common(a.top):
    print(f"{v / a.events:8.2f} {sz[kx] / 1024:5.1f} {100 * st[kx] / max(S, 1):6.1f}  {kx}")
if a.lines:
    lo, hi = map(int, a.lines.split("-"))
    for ln in sorted(per_line):
        v, b, smp, src = per_line[ln]
        if lo <= ln <= hi and (v or b): print(f"{ln:5d} {v / a.events:7.2f} {b:5d}B {100 * smp / max(S, 1):5.1f}%  {src.strip()[:96]}")
