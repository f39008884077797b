#!/usr/bin/env python
"""BUILD CONTAINER ONLY (needs /root/reference): how fast is the oracle PORT (oracle/factory_oracle.py) next to the
UNMODIFIED reference model (oracle/run_reference.py: /root/reference/src/manusim on the same SimPy layer) on one host
core, same seeds, same workload (bench.py's default: BASELINE config #4, full horizon)?  bench.py's CPU arm can only
time the port on the GPU box (the reference cannot travel); this calibration, committed as
profiles/cpu_calibration.json and quoted in every bench line's `cpu_baseline`, anchors its number to the real model
code.  Both use the reference's own convention: time of run_simulation() only (factory_sim.py:810-814).

usage: python tools/cpu_calibration.py [--reps 2]"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import bench
    from oracle import run_reference as rr
    from oracle.factory_oracle import FactoryOracle, SIMPY_ENGINE
    from oracle.variates import run_seeds

    spec = bench.workload_spec("jobshop20", None)
    seeds = run_seeds(bench.BASE_SEED, a.reps * len(spec["parts"]))
    rows = []
    k = 0
    for i, ((cfg, res, prod), policy, _, _) in enumerate(spec["parts"]):
        for _ in range(a.reps):
            seed = int(seeds[k]); k += 1
            cfg_l = dict(cfg, enable_event_logging=True)
            t0 = time.perf_counter()
            o = FactoryOracle(cfg_l, res, prod, seed, policy).run()
            t_port = time.perf_counter() - t0
            patches = ["P1"] if cfg.get("delivery_mode") == "onDue" else []
            harness = rr._make_harness(rr.load_reference()["FactorySimulation"], patches)
            sim = harness(dict(cfg, memory_size=400000), res, prod, print_mode="none", seed=seed)
            sim.reset_simulation(seed, None)
            t0 = time.perf_counter()
            sim.run_simulation()
            t_ref = time.perf_counter() - t0
            steps_ref = next(sim.env._eid) - len(sim.env._queue)
            assert steps_ref == o["steps"], (steps_ref, o["steps"])
            rows.append({"part": i, "delivery_mode": cfg.get("delivery_mode"), "seed": seed, "simpy_steps": o["steps"],
                         "port_s": t_port, "reference_s": t_ref})
            print(json.dumps(rows[-1]), flush=True)
    ev = sum(r["simpy_steps"] for r in rows)
    port = ev / sum(r["port_s"] for r in rows)
    ref = ev / sum(r["reference_s"] for r in rows)
    out = {"workload": spec["desc"], "simpy_layer": SIMPY_ENGINE, "replications": len(rows), "cores": 1,
           "port_events_per_s_per_core": port, "unmodified_reference_events_per_s_per_core": ref,
           "port_over_reference": port / ref,
           "how": "tools/cpu_calibration.py in the build container: same seeds, identical SimPy step counts; the "
                  "reference records its variates and trace through the golden-generation harness "
                  "(oracle/run_reference.py), which costs it a few percent",
           "rows": rows}
    (ROOT / "profiles" / "cpu_calibration.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({k_: out[k_] for k_ in ("port_events_per_s_per_core", "unmodified_reference_events_per_s_per_core",
                                            "port_over_reference")}))


if __name__ == "__main__":
    main()
