#!/usr/bin/env python
"""CPU (no GPU): parity at BASELINE.json's full horizons, per replication.  The kernel SOURCE runs on the CPU-fiber
emulation (tests/emu, test infrastructure) in Philox mode and records the variates it drew; the oracle replays them; the
complete record streams (ring, float32 time, float32 value, emission order) and the SimPy step counts must be equal.
This is the check that found the python-typed-clock case behind golden dbr_tiny_setup.

usage: python tools/full_horizon_check.py [--workloads jobshop20,toc_dbr,toc_dbr_shipped] [--reps N] [--seed S]
one JSON line per replication; exit status 1 if any differs."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "emu"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="jobshop20,toc_dbr,toc_dbr_shipped")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--seed", type=int, default=20261020)
    a = ap.parse_args()
    import numpy as np

    import bench
    import emu
    from manusim_b200 import compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    bad = 0
    for wl in a.workloads.split(","):
        spec = bench.workload_spec(wl, None)             # the bench's (= BASELINE's) full horizon
        for i, ((cfg, res, prod), policy, _, _) in enumerate(spec["parts"]):
            repair = spec.get("dbr_repair", False)
            t = compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250,
                                 max_demand_orders=250, fifo_capacity=256)
            eng = emu.EmuEngine(t)
            t0 = time.time()
            r = eng.run(a.reps, seed=a.seed + i, n_log_reps=a.reps, log_cap=1 << 21, record_tapes=1 << 21)
            emu_s = time.time() - t0
            cnt = r.rec["count"]
            assert cnt.max() < (1 << 21) and r.log_count.max() <= (1 << 21)
            for k in range(a.reps):
                tp = {s: r.rec[s][k, : cnt[k, j]] for j, s in enumerate("ABCD")}
                o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
                ring, tm, val, _ = decode_log(r.log[k], int(r.log_count[k]))
                same = (o["error"] is None and r.status[k] == 0 and int(r.n_events[k]) == o["steps"]
                        and len(ring) == len(o["ring"]) and np.array_equal(ring, o["ring"].astype(np.uint32))
                        and np.array_equal(tm, o["time"]) and np.array_equal(val, o["value"]))
                bad += not same
                print(json.dumps({"workload": wl, "part": i, "delivery_mode": cfg.get("delivery_mode"), "policy": policy,
                                  "run_until": cfg["run_until"], "replication": k, "simpy_steps": int(r.n_events[k]),
                                  "records": int(len(ring)), "identical_to_oracle_replay": bool(same),
                                  "emu_s_per_rep": round(emu_s / a.reps, 1)}), flush=True)
            eng.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
