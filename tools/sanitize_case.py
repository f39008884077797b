"""tiny run of every kernel variant for compute-sanitizer (memcheck / racecheck), one tool per gpurun call"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from manusim_b200 import BatchEngine, compile_scenario, scenarios
for policy, mode, repair in (("fifo", "asReady", False), ("max_priority", "onDue", False), ("edd", "instantly", False), ("dbr", "instantly", True)):
    if policy == "dbr":
        sim, res, prod, _ = scenarios.load_example("toc_dbr")
        cfg = dict(sim, run_until=301, warmup=50); cfg.pop("print_mode", None)
    else:
        cfg, res, prod = scenarios.jobshop20(run_until=200, warmup=30, mode=mode)
        cfg["log_queues"] = True
        for r in res.values(): r["tbf"] = {"dist": "expo", "params": [40]}
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair))
    r = eng.run(40, seed=1, n_log_reps=40, log_cap=512, record_tapes=4096)
    m = eng.reduce(r)
    torch.cuda.synchronize()
    h = eng.run_host(16, seed=2)
    print(policy, mode, "status ok:", bool((r.status == 0).all()), "events", int(r.n_events.sum()), "host", int(h.n_events.sum()))
    eng.close()
# the fast kernels (no queue records, no tape recording: constant-layout lean instantiations), every policy / delivery family
for policy, mode, repair in (("fifo", "asReady", False), ("fifo", "onDue", False), ("fifo", "instantly", False),
                             ("max_priority", "asReady", False), ("max_priority", "onDue", False), ("edd", "instantly", False),
                             ("dbr", "instantly", True), ("dbr", "onDue", True)):
    if policy == "dbr":
        sim, res, prod, _ = scenarios.load_example("toc_dbr")
        cfg = dict(sim, run_until=301, warmup=50, delivery_mode=mode); cfg.pop("print_mode", None)
    else:
        cfg, res, prod = scenarios.jobshop20(run_until=200, warmup=30, mode=mode)
        for r in res.values(): r["tbf"] = {"dist": "expo", "params": [40]}
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair))
    assert eng.info["const_layout"] == 1
    r = eng.run(40, seed=1, n_log_reps=40, log_cap=512)
    m = eng.reduce(r)
    torch.cuda.synchronize()
    print("fast", policy, mode, "status ok:", bool((r.status == 0).all()), "events", int(r.n_events.sum()))
    eng.close()
