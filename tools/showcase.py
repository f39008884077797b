#!/usr/bin/env python
"""BASELINE configs at full replication counts (run under torchrun for N>1):
  --config 4 : synthetic 20x10 job shop, 1 000 000 replications (even = asReady, odd = onDue), run_until=10000
  --config 5 : DBR demand-rate x buffer-size sweep, 256 scenarios x 31 250 replications = 8 000 000 (shortened
               horizon --run-until, default 2001; the shipped 200001 would be ~4.6e13 events)
  --config 3 : toc_dbr (repaired, P2+P3), 100 000 replications at the shipped horizon 200001
Prints one JSON line (rank 0): wall seconds, SimPy-equivalent events, events/s, replications/s, failures."""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--reps", type=int, default=0)
    ap.add_argument("--run-until", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0, help="replications per round of launches; 0 = whole waves of every part's kernel, about 3e5")
    a = ap.parse_args()
    import torch, torch.distributed as dist
    from manusim_b200 import BatchEngine, compile_scenario, scenarios
    from manusim_b200.engine import rep_keys
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    if a.config == 5:
        from manusim_b200.factory_sim import DBRSimulation
        from manusim_b200.sweep import SweepRunner
        sim, res, prod, _ = scenarios.load_example("toc_dbr")
        ru = a.run_until or 2001
        sim = dict(sim, run_until=ru, warmup=ru // 10)
        sim.pop("print_mode", None)
        rates = [10.0 / m for m in np.linspace(0.70, 1.00, 16)]
        bufs = list(range(50, 201, 10))
        sw = SweepRunner(DBRSimulation, {"simulation": sim, "resources": res, "products": prod},
                         {"products.*.demand.freq.params.0": rates, "products.*.shipping_buffer": bufs},
                         reps=a.reps or 31250, seed=2026,
                         sim_kwargs={"dbr_repair": True, "max_production_orders": 80, "max_demand_orders": 24})
        df = sw.run()
        wall = time.time() - t0
        n_total = len(sw.points) * sw.reps
        if rank == 0:
            lost = df[(df.variable == "lostSales")].groupby("scenario")["mean"].sum()
            print(json.dumps({"config": 5, "scenarios": len(sw.points), "replications": n_total, "gpus": world, "run_until": ru,
                              "horizon": ("the shipped toc_dbr horizon" if ru == 200001 else
                                          f"SHORTENED: run_until={ru} instead of the 200001 of the config-#3 tables"),
                              "launches": sw.launches, "reruns": sw.reruns, "streams": sw.streams,
                              "wall_s": wall, "events": sw.events, "events_per_s": sw.events / wall,
                              "replications_per_s": n_total / wall,
                              "lost_sales_per_run_min_max": [float(lost.min()), float(lost.max())]}))
    else:
        if a.config == 4:
            n_total = a.reps or 1_000_000
            ru = a.run_until or 10000
            parts = []
            for phase, mode, mdo in ((0, "asReady", 112), (1, "onDue", 128)):
                cfg, res, prod = scenarios.jobshop20(run_until=ru, warmup=ru // 10, mode=mode)
                parts.append((BatchEngine(compile_scenario(cfg, res, prod, max_production_orders=112, max_demand_orders=mdo)), phase, 2))
        else:
            n_total = a.reps or 100_000
            ru = a.run_until or 200001
            sim, res, prod, _ = scenarios.load_example("toc_dbr")
            cfg = dict(sim, run_until=ru, warmup=ru // 2)
            cfg.pop("print_mode", None)
            parts = [(BatchEngine(compile_scenario(cfg, res, prod, policy="dbr", dbr_repair=True, max_production_orders=80,
                                                   max_demand_orders=24)), 0, 1)]
        lo, hi = (n_total * rank) // world, (n_total * (rank + 1)) // world
        K = parts[0][0].K
        mom = torch.zeros((K, 3), dtype=torch.float64, device=dev)
        ev = torch.zeros((), dtype=torch.int64, device=dev)
        fails = torch.zeros((), dtype=torch.int64, device=dev)
        rerun = 0
        chunk = a.chunk
        if chunk <= 0:
            # kernel time is quantised in waves of resident replications: size a round as whole waves of every part
            import math
            whole = math.lcm(*[eng.resident_replications for eng, _, _ in parts]) * max(per for _, _, per in parts)
            chunk = whole * max(1, 300_000 // whole)
        for c0 in range(lo, hi, chunk):
            c1 = min(hi, c0 + chunk)
            for eng, phase, period in parts:
                g = np.arange(c0, c1, dtype=np.int64)
                g = g[g % period == phase]
                if len(g) == 0:
                    continue
                keys = rep_keys(20260101, g)
                r = eng.run(len(g), rep_seeds=keys)
                rerun += eng.retry_overflow(r, rep_seeds=keys)
                eng.reduce(r, out=mom)
                ev += r.n_events.sum()
                fails += (r.status != 0).sum()
        if world > 1:
            dist.all_reduce(mom); dist.all_reduce(ev); dist.all_reduce(fails)
        torch.cuda.synchronize()
        wall = time.time() - t0
        if rank == 0:
            m = mom.cpu().numpy()
            names = parts[0][0].tables.ring_names()
            util = {k: m[i, 0] / m[i, 2] for i, (v, k) in enumerate(names) if v == "utilization"}
            print(json.dumps({"config": a.config, "replications": n_total, "gpus": world, "run_until": ru, "wall_s": wall,
                              "events": int(ev.item()), "events_per_s": int(ev.item()) / wall,
                              "replications_per_s": n_total / wall, "failed": int(fails.item()), "rerun_this_rank": rerun,
                              "mean_utilization_min_max": [min(util.values()), max(util.values())]}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
