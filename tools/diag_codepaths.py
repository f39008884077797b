#!/usr/bin/env python
"""GPU diagnostic: how much faster does the simulation kernel run when whole code paths are never executed?  Same
20-resource job shop (asReady), same kernel binary; variants switch off what touches code, not what costs many
instructions: no records (warm-up = whole run: emit + flush never run), no breakdowns, constant demand gaps (no slow
variate path per order), all constant (no Philox at all).  A jump far beyond the saved instructions means the
instruction cache is the bound (profiles/README.md)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from manusim_b200 import BatchEngine, compile_scenario, scenarios

RU = int(sys.argv[1]) if len(sys.argv) > 1 else 2000


def run(name, cfg, res, prod):
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy="fifo", max_production_orders=112, max_demand_orders=112))
    n = 3 * eng.info["warps_per_block"] * eng.info["blocks_per_sm"] * eng.info["n_sms"]
    out = eng.alloc(n, n, 4096)
    ms = []
    for k in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(n, seed=5 + k, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ev = int(out.n_events.sum())
    rec = int(out.acc_cnt.sum(dtype=torch.int64))
    print(json.dumps({"variant": name, "events_per_s": ev / (min(ms[1:]) * 1e-3), "events_per_rep": ev / n, "records_per_event": rec / ev,
                      "failed": int((out.status != 0).sum()), "warps_per_sm": eng.info["warps_per_block"] * eng.info["blocks_per_sm"]}), flush=True)
    eng.close()


cfg, res, prod = scenarios.jobshop20(run_until=RU, warmup=RU // 10, mode="asReady")
run("base", cfg, res, prod)
run("no_records", dict(cfg, warmup=RU - 1), res, prod)
import copy

r2 = copy.deepcopy(res)
for r in r2.values():
    r.pop("tbf", None); r.pop("ttr", None)
run("no_breakdowns", cfg, r2, prod)
p2 = copy.deepcopy(prod)
for p in p2.values():
    p["demand"]["freq"] = {"dist": "constant", "params": [p["demand"]["freq"]["params"][0]]}
run("constant_demand_gaps", cfg, res, p2)
run("no_records_no_breakdowns_constant_gaps", dict(cfg, warmup=RU - 1), r2, p2)
r3, p3 = scenarios.make_deterministic(r2, p2)
run("all_constant", cfg, r3, p3)
run("all_constant_no_records", dict(cfg, warmup=RU - 1), r3, p3)
