#!/usr/bin/env python
"""CPU (no GPU): long-horizon parity sweep over scenario kinds the BASELINE configs do not reach (simple_scheduler to
t = 20 000, onDue and EDD variants, DBR x onDue, a python-typed-clock stress case, job shop with lots + queue records +
breakdowns).  Kernel source on the CPU-fiber emulation (Philox + tape recording) against oracle replays; one JSON line per
replication.  This sweep found the two rare cases behind goldens dbr_tiny_setup and ss_zero_ops; the stress case
`dbr_typed_long` is the python-typed-clock stress case that golden dbr_f64_order was cut from (round 2: identical)."""
import sys, time, json, copy
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tests' / 'emu')); sys.path.insert(0, str(ROOT))
import numpy as np, emu
from manusim_b200 import compile_scenario, scenarios
from manusim_b200.engine import decode_log
from oracle.factory_oracle import run_oracle
cases = []
sim, res, prod, _ = scenarios.load_example("simple_scheduler"); sim.pop("print_mode", None)
cases.append(("ss_long", dict(sim, run_until=20000, warmup=1000), res, prod, "max_priority", False))
cases.append(("ss_ondue_long", dict(sim, run_until=10000, warmup=500, delivery_mode="onDue"), res, prod, "max_priority", False))
cases.append(("ss_edd_long", dict(sim, run_until=8000, warmup=500), res, prod, "edd", False))
simd, resd, prodd, _ = scenarios.load_example("toc_dbr"); simd.pop("print_mode", None)
cases.append(("dbr_ondue_long", dict(simd, run_until=60001, warmup=20000, delivery_mode="onDue"), resd, prodd, "dbr", True))
rest = copy.deepcopy(resd)
for i, r in enumerate(rest.values()):
    if i % 2 == 0: r["setup"] = {"dist": "constant", "params": [0, 0]}
cases.append(("dbr_typed_long", dict(simd, run_until=100001, warmup=50000, scheduler_interval=24, cb_target_level=120), rest, prodd, "dbr", True))
cfg, r20, p20 = scenarios.jobshop20(run_until=10000, warmup=1000, mode="instantly")
cfg["log_queues"] = True
for i, r in enumerate(r20.values()):
    r["tbf"] = {"dist": "expo", "params": [400 + 11 * i]}; r["ttr"] = {"dist": "gamma", "params": [4, 3]}
for j, p_ in enumerate(p20.values()):
    p_["demand"]["quantity"] = {"dist": "uniform", "params": [2, 0.5]}; p_["demand"]["freq"] = {"dist": "erlang", "params": [24 + j, 12]}
cases.append(("js20_lots_queues_long", cfg, r20, p20, "fifo", False))
bad = 0
for name, cfg, res, prod, policy, repair in cases:
    t = compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250, max_demand_orders=250, fifo_capacity=256)
    eng = emu.EmuEngine(t)
    n = 2
    t0 = time.time()
    r = eng.run(n, seed=9090, n_log_reps=n, log_cap=1 << 21, record_tapes=1 << 21)
    es = time.time() - t0
    cnt = r.rec["count"]
    for k in range(n):
        tp = {s: r.rec[s][k, : cnt[k, j]] for j, s in enumerate("ABCD")}
        o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
        ring, tm, val, _ = decode_log(r.log[k], int(r.log_count[k]))
        m = min(len(ring), len(o["ring"]))
        neq = (ring[:m] != o["ring"][:m].astype(np.uint32)) | (tm[:m] != o["time"][:m]) | (val[:m] != o["value"][:m])
        same = o["error"] is None and r.status[k] == 0 and int(r.n_events[k]) == o["steps"] and len(ring) == len(o["ring"]) and not neq.any()
        bad += not same
        print(json.dumps({"case": name, "rep": k, "status": int(r.status[k]), "steps": int(r.n_events[k]), "oracle_steps": o["steps"], "records": len(ring), "identical": bool(same), "n_diff": int(neq.sum()), "first_diff": int(np.argmax(neq)) if neq.any() else None, "oracle_error": o["error"], "emu_s": round(es/n,1)}), flush=True)
print("bad", bad)
