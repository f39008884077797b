#!/usr/bin/env python
"""Offline estimate of warp instructions per SimPy event for a build of the simulation kernel -- no GPU needed.

    estimate = sum over SASS instructions of  (expected executions of the instruction per replication warp)
               / SimPy events of the run

* SASS side: `nvdisasm -gi` of the cubin inside manusim_b200/libmsim.so gives every instruction of
  sim_kernel<POLICY,RNG> its source line and inlining chain (compiled with -lineinfo).
* execution side: the SAME source runs on the host-side SIMT emulation (tests/emu, test infrastructure) built with
  gcov; a representative workload gives per-line execution counts.  An instruction whose chain is
  l0 (kernel) <- l1 <- ... <- lk (innermost) is charged  exec(l0) * prod_j exec(lj) / entries(function of lj):
  exact for code inlined at one site, an average over call sites otherwise.  Kernel-function lines that all 32 lanes
  execute are divided by 32 (one warp instruction), lane-0 code is not.
It is a model: both branches of a predicated/if-converted region count, instruction replays and the scheduling of
independent lines into each other are ignored.  It is calibrated against the one number ncu measured on a B200 for
the default kernel (86 warp instructions per SimPy event on jobshop20, profiles/README.md) and is meant for RELATIVE
comparisons between kernel variants before GPU minutes are spent on them.

usage: python tools/instr_model.py [--layout soa|aos] [--workload jobshop20|simple_det|toc_dbr] [--run-until N] [--top K]
                                   [--part i] [--defines MSIM_X_...=1,... --lib manusim_b200/libmsim_<tag>.so]
Track record: it mis-predicted the AoS slab (a code-SIZE effect: predicted ~0, measured +7 %); for changes that remove
executed instructions from straight-line code (e.g. the onDue timer cache) it is the right tool.
"""
from __future__ import annotations

import argparse
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))
BUILD = ROOT / "tests" / "emu" / "_build"

SRC = {"soa": "msim_kernel.cu", "aos": "msim_kernel_aos.cu"}
POLICY_ID = {"fifo": 0, "max_priority": 1, "dbr": 2, "edd": 3}


# ----------------------------------------------------------------------------------------------- execution counts
RUNNER = r"""
import json, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {emu!r})
import numpy as np
import emu, bench
from manusim_b200 import compile_scenario
spec = bench.workload_spec({workload!r}, {run_until!r})
events = 0
for i, ((cfg, res, prod), policy, phase, period) in enumerate(spec["parts"]):
    mpo, mdo = spec.get("pools", [(0, 0)] * len(spec["parts"]))[i]
    t = compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False),
                         max_production_orders=mpo, max_demand_orders=mdo)
    if {part!r} is not None and i != {part!r}:
        continue
    eng = emu.EmuEngine(t, "cov", layout={layout!r}, defines={defines!r})
    r = eng.run({reps}, seed=20260101 + i, n_log_reps={reps}, log_cap=4096)
    assert (r.status == 0).all(), r.status
    events += int(r.n_events.sum())
    eng.close()
print(json.dumps({{"events": events, "policy": spec["parts"][0][1]}}))
"""


def run_coverage(layout, workload, run_until, reps, defines=(), part=None):
    import emu

    emu.build("cov", extra_defines=defines)
    for f in BUILD.glob("*.gcda"):
        f.unlink()
    code = RUNNER.format(root=str(ROOT), emu=str(ROOT / "tests" / "emu"), workload=workload, run_until=run_until,
                         layout=layout, reps=reps, defines=tuple(defines), part=part)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=3600,
                       env={**__import__("os").environ, "MSIM_EMU_THREADS": "1"})
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])
    info = json.loads(r.stdout.strip().splitlines()[-1])
    stem = Path(SRC[layout]).stem + "_" + emu._tag("cov", tuple(defines))
    g = subprocess.run(["gcov", "--json-format", "--stdout", str(BUILD / f"{stem}.gcda")], capture_output=True,
                       text=True, cwd=str(BUILD))
    if g.returncode != 0:
        raise RuntimeError(g.stderr[-2000:])
    data = json.loads(g.stdout)
    lines = defaultdict(lambda: defaultdict(int))      # file -> line -> count
    funcs = defaultdict(list)                          # file -> [(start, end)]
    for f in data["files"]:
        name = Path(f["file"]).name
        for ln in f["lines"]:
            lines[name][ln["line_number"]] += ln["count"]
        seen = set()
        for fn in f["functions"]:
            key = (fn["start_line"], fn["end_line"])
            if key not in seen:
                seen.add(key)
                funcs[name].append(key)
    return info, lines, funcs


# ----------------------------------------------------------------------------------------------- SASS with lines
def sass_chains(layout, policy, rng=0, lib=None):
    tmp = BUILD / "cubin"
    tmp.mkdir(exist_ok=True)
    for f in tmp.glob("*.cubin"):
        f.unlink()
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve() if lib else ROOT / "manusim_b200" / "libmsim.so")], cwd=str(tmp), check=True,
                   capture_output=True)
    cubin = tmp / (Path(SRC[layout]).stem + ".sm_100a.cubin")
    txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
    # mangled kernel name: the AoS kernels carry a third template argument (OPTS; 3 = the general kernel)
    want = f"sim_kernelILi{POLICY_ID[policy]}ELi{rng}E" + ("Li3EEEv" if layout == "aos" else "EEv")
    out = []                      # (opcode, chain) ; chain = [(file, line)] innermost first
    active = False
    pending = []
    cur = None
    tag = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);")
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            active = want in line
            pending, cur = [], None
            continue
        if not active:
            continue
        m = tag.search(line)
        if m:
            pending.append((Path(m.group(1)).name, int(m.group(2))))
            if "inlined at" not in line:          # last tag of a chain
                cur, pending = pending, []
            continue
        m = ins.match(line)
        if m and cur is not None:
            out.append((m.group(1).split()[0] if not m.group(1).startswith("@") else m.group(1).split()[1], cur))
    return out


def function_ranges(src_path):
    """(start, end) line ranges of the __device__/__global__ function bodies of a source file, by brace matching
    (gcov reports no function records for always_inline functions)."""
    text = src_path.read_text().splitlines()
    out = []
    i = 0
    while i < len(text):
        if re.search(r"__(device|global)__", text[i]) and not text[i].lstrip().startswith("//"):
            j = i
            while j < len(text) and "{" not in text[j] and ";" not in text[j]:
                j += 1
            if j < len(text) and "{" in text[j]:
                d = 0
                k = j
                while k < len(text):
                    d += text[k].count("{") - text[k].count("}")
                    if d <= 0:
                        break
                    k += 1
                out.append((i + 1, k + 1))
                i = j + 1            # nested lambdas / inner functions do not occur; continue inside for safety
                continue
        i += 1
    return out


def function_names(src_path, ranges):
    text = src_path.read_text().splitlines()
    out = {}
    for s_, e_ in ranges:
        head = " ".join(text[s_ - 1:s_ + 2])
        m = re.search(r"(\w+)\s*\(", re.sub(r"__\w+__(\(.*?\))?", "", head))
        out[(s_, e_)] = m.group(1) if m else f"line{s_}"
    return out


def lane0_lines(src_path):
    """Line numbers of sim_kernel() that only lane 0 executes (brace-matched `if (lane == 0)` blocks), and the
    line range of the kernel function."""
    text = src_path.read_text().splitlines()
    start = next(i for i, l in enumerate(text) if "sim_kernel(const __grid_constant__ KParams p)" in l)
    depth, end = 0, None
    for i in range(start, len(text)):
        depth += text[i].count("{") - text[i].count("}")
        if depth == 0 and i > start:
            end = i
            break
    lane0 = set()
    cond = set()                     # lines that hold the `if (lane == 0)` itself: all 32 lanes evaluate it
    i = start
    while i <= end:
        if re.search(r"if \(lane == 0\)", text[i]):
            cond.add(i + 1)
            d = 0
            j = i
            opened = False
            while j <= end:
                d += text[j].count("{") - text[j].count("}")
                opened |= "{" in text[j]
                lane0.add(j + 1)
                if (opened and d <= 0) or (not opened and text[j].rstrip().endswith(";")):
                    break
                j += 1
            i = j + 1
        else:
            i += 1
    return lane0 - cond, (start + 1, end + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="soa", choices=["soa", "aos"])
    ap.add_argument("--workload", default="jobshop20")
    ap.add_argument("--run-until", type=int, default=3000)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--json", default=None)
    ap.add_argument("--defines", default="", help="comma-separated experiment switches, e.g. MSIM_X_DUECACHE=1 "
                                                   "(needs the matching nvcc build: --lib manusim_b200/libmsim_<tag>.so)")
    ap.add_argument("--lib", default=None)
    ap.add_argument("--part", type=int, default=None, help="only this part of the workload (jobshop20: 0 asReady, 1 onDue)")
    a = ap.parse_args()
    defines = tuple(d for d in a.defines.split(",") if d)

    info, lines, funcs = run_coverage(a.layout, a.workload, a.run_until, a.reps, defines, a.part)
    src = SRC[a.layout]
    funcs = {name: function_ranges(ROOT / "manusim_b200" / "csrc" / name) for name in (src, "philox.cuh")}
    chains = sass_chains(a.layout, info["policy"], lib=a.lib)
    lane0, (k0, k1) = lane0_lines(ROOT / "manusim_b200" / "csrc" / src)

    def entries(fname, line):
        """executions of the first counted line of the function that contains `line`"""
        best = None
        for s, e in funcs.get(fname, []):
            if s <= line <= e and (best is None or s > best[0]):
                best = (s, e)
        if best is None:
            return None
        for ln in range(best[0], best[1] + 1):
            c = lines[fname].get(ln)
            if c:
                return c
        return 0

    names = {name: function_names(ROOT / "manusim_b200" / "csrc" / name, funcs[name]) for name in funcs}

    def func_name(fname, line):
        best = None
        for (s_, e_), nm in names.get(fname, {}).items():
            if s_ <= line <= e_ and (best is None or s_ > best[0]):
                best = (s_, nm)
        return f"{fname}:{best[1]}" if best else f"{fname}:?"

    def count(fname, line):
        """gcov count of a line; a call line of an always_inline function often has no counter of its own: take the
        nearest instrumented line after it (the statement following the call), else before it"""
        tab = lines[fname]
        if line in tab:
            return float(tab[line])
        for d in (1, 2, 3, -1, -2, -3):
            if line + d in tab:
                return float(tab[line + d])
        return 0.0

    total = 0.0
    warm = 0                    # static instructions executed more than once per 1000 SimPy events
    warm_by_func = defaultdict(int)
    by_func = defaultdict(float)
    by_line = defaultdict(float)
    by_op = defaultdict(float)
    static = len(chains)
    unknown = 0
    for op, chain in chains:
        outer_f, outer_l = chain[-1]
        if outer_f not in lines:
            unknown += 1
            continue
        ex = count(outer_f, outer_l)
        if outer_f == src and k0 <= outer_l <= k1 and outer_l not in lane0:
            ex /= 32.0                                   # all lanes run it: one warp instruction
        for f, l in reversed(chain[:-1]):                # walk inwards
            if f not in lines:                           # CUDA headers (intrinsics): executed with their caller
                continue
            ent = entries(f, l)
            if not ent:
                ex = 0.0
                break
            ex *= count(f, l) / ent          # (> 1 for lines inside loops of the inlined function)
        total += ex
        if ex / info["events"] > 1e-3:
            warm += 1
            warm_by_func[func_name(*next(((f, l) for f, l in chain if f in lines), chain[-1]))] += 1
        inner = next(((f, l) for f, l in chain if f in lines), chain[-1])      # innermost line of OUR sources
        by_line[inner] += ex
        by_func[func_name(inner[0], inner[1])] += ex
        by_op[op] += ex
    ev = info["events"]
    res = {"layout": a.layout, "workload": a.workload, "run_until": a.run_until, "events": ev,
           "static_instructions": static, "est_warp_instr_per_event": total / ev,
           "unattributed_static": unknown, "warm_static_instructions": warm, "warm_code_kb": warm * 16 / 1024,
           "by_function": {k: v / ev for k, v in sorted(by_func.items(), key=lambda kv: -kv[1])}}
    print(json.dumps(res))
    print(f"\ntop source lines ({src}) by estimated warp instructions per event")
    for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1])[: a.top]:
        print(f"  {v / ev:7.3f}  {f}:{l}")
    print("\nby function (innermost source function of each instruction)")
    for fn, v in sorted(by_func.items(), key=lambda kv: -kv[1])[:20]:
        print(f"  {v / ev:7.3f}  {fn}")
    print("\nwarm static instructions by function (x16 B)")
    for fn, v in sorted(warm_by_func.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {v:6d}  {fn}")
    print("\nby opcode")
    for op, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:15]:
        print(f"  {v / ev:7.3f}  {op}")
    if a.json:
        Path(a.json).write_text(json.dumps({**res, "by_line": {f"{f}:{l}": v / ev for (f, l), v in by_line.items()},
                                            "by_opcode": {k: v / ev for k, v in by_op.items()}}, indent=1))


if __name__ == "__main__":
    main()
