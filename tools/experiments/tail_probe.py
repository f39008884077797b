#!/usr/bin/env python
"""GPU: kernel time of ONE workload part against the number of replications per launch (in units of resident
replications), to separate the per-wave cost from the per-launch tail.  One JSON line per (part, n)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np
import torch
import bench
from manusim_b200 import BatchEngine, compile_scenario

spec = bench.workload_spec("jobshop20", int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
dev = torch.device("cuda", 0)
for i, ((cfg, res, prod), policy, _, _) in enumerate(spec["parts"]):
    mpo, mdo = spec["pools"][i]
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, max_production_orders=mpo, max_demand_orders=mdo))
    resident = eng.info["warps_per_block"] * eng.info["blocks_per_sm"] * eng.info["n_sms"]
    for waves in (1.0, 1.5, 2.0, 3.0, 4.0, 6.0):
        n = int(round(waves * resident))
        out = eng.alloc(n, 0, 0)
        best = None
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            eng.run(n, seed=7 + rep, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        ev = int(out.n_events.sum().item())
        print(json.dumps({"part": i, "delivery_mode": cfg["delivery_mode"], "warps_per_sm": resident // eng.info["n_sms"],
                          "waves": waves, "replications": n, "ms": best, "events_per_s": ev / best * 1e3,
                          "failed": int((out.status != 0).sum().item())}), flush=True)
    eng.close()
