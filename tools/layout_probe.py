#!/usr/bin/env python
"""GPU: smoke() for both slab layouts, then one short timing of each on the bench workload (jobshop20, shortened).
Not a bench line -- a quick A/B of the two slab layouts (array-of-structures = default, structure-of-arrays =
MSIM_SLAB_LAYOUT=soa) on identical trajectories.
usage: python tools/layout_probe.py [run_until] [replications] [workload] [nosmoke]"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ru = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5920
    workload = sys.argv[3] if len(sys.argv) > 3 else "jobshop20"
    smoke = not (len(sys.argv) > 4 and sys.argv[4] == "nosmoke")
    import torch

    import __graft_entry__ as g
    import bench
    from manusim_b200 import BatchEngine, compile_scenario

    out = {}
    for layout in ("soa", "aos"):
        os.environ["MSIM_SLAB_LAYOUT"] = layout
        t0 = time.time()
        if smoke:
            g.smoke()
        out[f"smoke_{layout}_s"] = round(time.time() - t0, 2)
        spec = bench.workload_spec(workload, ru)
        (cfg, res, prod), policy, _, _ = spec["parts"][-1]
        mpo, mdo = spec["pools"][-1]
        eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False),
                                           max_production_orders=mpo, max_demand_orders=mdo))
        r = eng.alloc(n, 0, 0)
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(n, seed=7 + rep, out=r)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        ev = int(r.n_events.sum())
        out[layout] = {"ms": ms, "events": ev, "events_per_s": ev / (ms * 1e-3), "failed": int((r.status != 0).sum()),
                       "info": {k: eng.info[k] for k in ("smem_per_replication", "warps_per_block", "blocks_per_sm",
                                                         "regs_per_thread")}}
        print(json.dumps({layout: out[layout]}), flush=True)
    os.environ.pop("MSIM_SLAB_LAYOUT", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
