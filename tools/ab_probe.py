#!/usr/bin/env python
"""A/B harness for kernel experiments: build variants on CPU, check them bit for bit on the CPU-fiber emulation, then
time all of them in ONE GPU call on identical trajectories.

    python tools/ab_probe.py build                 # nvcc: manusim_b200/libmsim_<tag>.so for every variant (no GPU)
    python tools/ab_probe.py verify                # emulator: reference goldens + fuzz seeds per variant (no GPU)
    python tools/ab_probe.py static                # SASS instruction counts per variant (no GPU)
    python tools/ab_probe.py run [--run-until N] [--reps N] [--rounds K] [--workloads jobshop20,toc_dbr]   # GPU

Variants are -D switches of csrc/msim_kernel.cu (none = the default build).  `run` executes every variant in its own process (the library is chosen at load time with MSIM_LIB),
K rounds interleaved so that clock drift shows up as spread instead of bias, after a smoke() parity check per variant.
Output: one JSON line per (round, variant, workload part) and a summary table; not a bench line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

# round-2 history (profiles/r02_ab_probe.jsonl): tailpush +10 %, lean +15..23 %, lb4 +12 % where shared memory limits the
# handle to 4 CTAs per SM -> folded into the default build; flushgrp -5 %, duecache -5 %, ropecache 0, nto 0 -> removed.
# second batch (profiles/r02_ab_diet_*.jsonl, patch kept as tools/experiments/r02_code_diet.patch): delivery mode as a
# template parameter + one flush site + cold draws out of line -5 %, calendar pop through the attention word +3 % on top of
# that, per-stripe earliest onDue timer +2 % -- none beat the plain build, whose 36 KB hot set sits right at the
# instruction-cache capacity: the response to small code changes is dominated by layout, not by instruction count.
# third batch (profiles/r02_ab_variants_*.jsonl), each change switchable on its own: delivery mode as a template
# parameter +7..8 % (job shop) -> adopted for every lean kernel; calendar pop through the attention word + single flush
# site: -4 % on the FIFO kernels, +9 % max-priority, +6 % DBR -> adopted per kernel family (LOOP2 in msim_kernel.cu);
# cold draws out of line: +11 % code (the call sequences outweigh the inlined constant test) -> dropped.
# fourth batch (profiles/r02_ab_variants_k_*.jsonl): sequence numbers filled in by the flush + one staging check per dispatch
# (emit2), `instantly` as its own kernel (mode3), rolled Philox rounds -- together +14 % on the max-priority kernels, -10 % on
# the FIFO kernels (rolled alone +3 % there) -> per-family compile-time choice (struct Tune in msim_kernel.cu).
# fifth batch (profiles/r02_ab_variants_n_*.jsonl, tools/experiments/r02_flush2_servefg.patch): log-ring position without the
# 64-bit modulo + rolled flush loop (-5 % code) and the finished-goods getter service out of line (-1.6 % code): -9 % / -5 %
# alone, +1 % together on the FIFO kernels -- smaller is not faster on this kernel; left out.
VARIANTS = {
}
RESTRICTED = set()
# variants of the DEFAULT library selected by environment (no rebuild)
ENV_VARIANTS = {"general": {"MSIM_NO_LEAN": "1"}}

WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
import torch
import bench
from manusim_b200 import BatchEngine, compile_scenario
if {smoke!r}:
    import __graft_entry__ as g
    g.smoke()
for wl in {workloads!r}:
    spec = bench.workload_spec(wl, {run_until!r} if wl == "jobshop20" else {run_until_dbr!r})
    for i, ((cfg, res, prod), policy, phase, period) in enumerate(spec["parts"]):
        mpo, mdo = spec.get("pools", [(0, 0)] * len(spec["parts"]))[i]
        eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False),
                                           max_production_orders=mpo, max_demand_orders=mdo))
        n = {reps!r}
        r = eng.alloc(n, n, 4096)                       # log ring like bench.py
        ms = []
        for k in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(n, seed=11 + k, out=r)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ev = int(r.n_events.sum())                       # events of the last launch
        print(json.dumps({{"variant": {tag!r}, "round": {rnd!r}, "workload": wl, "part": i, "ms": ms[1:],
                          "events": ev, "events_per_s": ev / (ms[-1] * 1e-3), "failed": int((r.status != 0).sum()),
                          "warps_per_sm": eng.info["warps_per_block"] * eng.info["blocks_per_sm"],
                          "regs": eng.info["regs_per_thread"]}}), flush=True)
        eng.close()
"""


def cmd_build(_):
    from manusim_b200.build import build_lib

    print(build_lib(force=True))
    for tag, defs in VARIANTS.items():
        print(build_lib(force=True, defines=defs, tag=tag))


def cmd_static(_):
    import re

    from manusim_b200.build import LIB, variant_path

    for tag in ["default"] + list(VARIANTS):
        lib = LIB if tag == "default" else variant_path(tag)
        if not lib.exists():
            continue
        txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
        name, cnt = None, {}
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = m.group(1)
                cnt[name] = 0
            elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", line):
                cnt[name] += 1
        row = {k.replace("_ZN4msim10sim_kernelILi", "<").replace("EEEvNS_7KParamsE", ">").replace("ELi", ","): v
               for k, v in cnt.items() if "10sim_kernel" in k}
        print(tag, json.dumps(row))


def cmd_verify(_):
    sys.path.insert(0, str(ROOT / "tests"))
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import numpy as np
    from conftest import ERROR_NAMES, GOLDEN_NAMES, load_golden

    import emu
    from manusim_b200 import compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle
    from test_gpu_fuzz import random_scenario

    pol = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}
    for tag, defs in VARIANTS.items():
        bad = 0
        for name in GOLDEN_NAMES + ERROR_NAMES:
            if tag in RESTRICTED and name == "ss_queues":
                continue
            g = load_golden(name)
            scn = g["scenario"]
            t = compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy=pol[scn["policy"]],
                                 dbr_repair="P3" in scn["patches"])
            ref = emu.EmuEngine(t)
            eng = emu.EmuEngine(t, defines=defs)
            tapes = [{k: g["tape_" + k] for k in "ABCD"}] * 2
            r0 = ref.run(2, tapes=tapes, n_log_reps=2, log_cap=len(g["ring"]) + 64)
            r = eng.run(2, tapes=tapes, n_log_reps=2, log_cap=len(g["ring"]) + 64)
            ring, tm, val, _ = decode_log(r.log[1], int(r.log_count[1]))
            ok = (np.array_equal(ring, g["ring"].astype(np.uint32)) and np.array_equal(tm, g["time"])
                  and np.array_equal(val, g["value"]) and np.array_equal(r.status, r0.status)
                  and np.array_equal(r.n_events, r0.n_events) and np.array_equal(r.n_timeouts, r0.n_timeouts)
                  and np.array_equal(r.acc_cnt, r0.acc_cnt) and np.array_equal(r.acc_sum, r0.acc_sum))
            if name in GOLDEN_NAMES:
                ok = ok and (r.status == 0).all() and (r.n_events == g["steps"]).all()
            bad += not ok
        for seed in () if tag in RESTRICTED else range(0, 48, 3):
            cfg, res, prod, policy, repair = random_scenario(seed)
            t = compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250,
                                 max_demand_orders=250, fifo_capacity=256)
            eng = emu.EmuEngine(t, defines=defs)
            r = eng.run(2, seed=1000 + seed, n_log_reps=2, log_cap=1 << 16, record_tapes=1 << 14)
            cnt = r.rec["count"]
            for i in range(2):
                tp = {k: r.rec[k][i, : cnt[i, j]] for j, k in enumerate("ABCD")}
                o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
                ring, tm, val, _ = decode_log(r.log[i], int(r.log_count[i]))
                ok = (np.array_equal(ring, o["ring"].astype(np.uint32)) and np.array_equal(tm, o["time"])
                      and np.array_equal(val, o["value"]))
                if o["error"] is None:
                    ok = ok and r.status[i] == 0 and int(r.n_events[i]) == o["steps"]
                bad += not ok
        print(json.dumps({"variant": tag, "defines": defs, "mismatches": bad}), flush=True)


def cmd_run(a):
    from manusim_b200.build import variant_path

    tags = [("default", {})] + list(ENV_VARIANTS.items())
    for tag in VARIANTS:
        if variant_path(tag).exists():
            tags.append((tag, {"MSIM_LIB": str(variant_path(tag))}))
    rows = []
    for rnd in range(a.rounds):
        for tag, env in tags:
            code = WORKER.format(root=str(ROOT), smoke=(rnd == 0), workloads=a.workloads.split(","), run_until=a.run_until,
                                 run_until_dbr=a.run_until_dbr, reps=a.reps, tag=tag, rnd=rnd)
            r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True,
                               timeout=a.timeout)
            if r.returncode != 0:
                print(json.dumps({"variant": tag, "round": rnd, "error": r.stderr[-400:]}), flush=True)
                continue
            for line in r.stdout.splitlines():
                if line.startswith("{"):
                    print(line, flush=True)
                    rows.append(json.loads(line))
    best = {}
    for r in rows:
        k = (r["variant"], r["workload"], r["part"])
        best[k] = max(best.get(k, 0.0), r["events_per_s"])
    print("\nbest events/s per variant (relative to default)")
    for (v, w, p_), e in sorted(best.items(), key=lambda kv: (kv[0][1], kv[0][2], kv[0][0])):
        base = best.get(("default", w, p_), e)
        print(f"  {w}[{p_}] {v:8s} {e:.4e}  {100 * (e / base - 1):+.1f} %")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "verify", "static", "run"])
    ap.add_argument("--run-until", type=int, default=2000)
    ap.add_argument("--run-until-dbr", type=int, default=6001)
    ap.add_argument("--reps", type=int, default=5920)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--workloads", default="jobshop20,toc_dbr")
    ap.add_argument("--timeout", type=int, default=120)
    a = ap.parse_args()
    {"build": cmd_build, "verify": cmd_verify, "static": cmd_static, "run": cmd_run}[a.cmd](a)


if __name__ == "__main__":
    main()
