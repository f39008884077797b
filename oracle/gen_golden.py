"""Generate tests/golden/* by executing the UNMODIFIED reference (/root/reference) on oracle/minisimpy.py.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python -m oracle.gen_golden            # all scenarios
    python -m oracle.gen_golden ss_cfg1    # one

Each ``tests/golden/<name>.npz`` holds, for one replication of one scenario: the ordered record stream
(ring u16, time f32, value f32), the four variate tapes (A,B,D f32; C f64), the SimPy step count, the patch
set, and the JSON-encoded scenario (config/resources/products/policy/seed) so that tests can rebuild the
exact inputs on a box that has no /root/reference.
"""
from __future__ import annotations

import copy
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from manusim_b200 import scenarios as sc  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def _example(name):
    sim, res, prod, exp = rr.load_example(name)
    return copy.deepcopy(sim), copy.deepcopy(res), copy.deepcopy(prod), exp


def scenario_table():
    """name -> dict(cfg, resources, products, policy, seed, patches)."""
    out = {}
    cfg, res, prod = sc.toy2()
    out["toy2"] = dict(cfg=cfg, resources=res, products=prod, policy="base", seed=1, patches=[])
    cfg, res, prod = sc.toy2(warmup=25, mode="instantly", run_until=150)
    out["toy2_instantly"] = dict(cfg=cfg, resources=res, products=prod, policy="base", seed=1, patches=[])

    sim, res, prod, _ = _example("simple_scheduler")
    sim["memory_size"] = 100000
    sim.pop("print_mode", None)
    out["ss_cfg1"] = dict(cfg=sim, resources=res, products=prod, policy="simple_scheduler", seed=54907, patches=[])

    sim2 = dict(sim, run_until=400)
    dres, dprod = sc.make_deterministic(res, prod)
    out["ss_det"] = dict(cfg=sim2, resources=dres, products=dprod, policy="simple_scheduler", seed=3, patches=[])

    sim3 = dict(sim, run_until=400, warmup=30, delivery_mode="instantly")
    out["ss_instantly"] = dict(cfg=sim3, resources=res, products=prod, policy="base", seed=280679, patches=[])

    # lots: demand quantity uniform -> round(x, 0) in {2,3,4}; batch sums with count > 1
    lot_prod = copy.deepcopy(prod)
    for p in lot_prod.values():
        p["demand"]["quantity"] = {"dist": "uniform", "params": [3, 0.5]}
        p["demand"]["freq"] = {"dist": "expo", "params": [34]}
        p["demand"]["duedate"] = {"dist": "normal", "params": [60, 25]}
    sim4 = dict(sim, run_until=500, warmup=40)
    out["ss_lots"] = dict(cfg=sim4, resources=res, products=lot_prod, policy="base", seed=91421, patches=[])

    cfg, res20, prod20 = sc.jobshop20(run_until=1200, warmup=150, mode="asReady")
    out["js20_asready"] = dict(cfg=cfg, resources=res20, products=prod20, policy="base", seed=20260101, patches=[])
    cfg, res20, prod20 = sc.jobshop20(run_until=1200, warmup=150, mode="onDue")
    out["js20_ondue"] = dict(cfg=cfg, resources=res20, products=prod20, policy="base", seed=20260102, patches=["P1"])
    # frequent breakdowns so TTR/TBF logic is exercised inside a short run
    cfg, res20b, prod20b = sc.jobshop20(run_until=1200, warmup=150, mode="asReady")
    for i, r in enumerate(res20b.values()):
        r["tbf"] = {"dist": "expo", "params": [60 + 7 * i]}
        r["ttr"] = {"dist": "normal", "params": [3, 2]}
    out["js20_breakdowns"] = dict(cfg=cfg, resources=res20b, products=prod20b, policy="base", seed=424242, patches=[])

    sim, res, prod, _ = _example("toc_dbr")
    sim.pop("print_mode", None)
    sim.update(run_until=3001, warmup=600, memory_size=200000)
    out["dbr_shipped"] = dict(cfg=sim, resources=res, products=prod, policy="dbr", seed=54907, patches=["P2"])
    out["dbr_repaired"] = dict(cfg=dict(sim), resources=res, products=prod, policy="dbr", seed=54907,
                               patches=["P2", "P3"])
    # python-typed clock excursions: zero setups on every other resource make non-CCR orders
    # released at int-typed tick instants start processing with an int clock + float batch time -> float64 instants
    res_t = copy.deepcopy(res)
    for i, r in enumerate(res_t.values()):
        if i % 2 == 0:
            r["setup"] = {"dist": "constant", "params": [0, 0]}
    sim_t = dict(sim, run_until=2501, warmup=400, scheduler_interval=24, cb_target_level=120)
    out["dbr_typed"] = dict(cfg=sim_t, resources=res_t, products=prod, policy="dbr", seed=91421, patches=["P2", "P3"])
    # DBR release limits (toc_dbr/simulation.py:245-262): CCR load budget per tick and at most 3 CCR releases per tick
    sim_l = dict(sim, run_until=2001, warmup=300, ccr_release_limit=0.6, order_release_limit=3)
    out["dbr_limits"] = dict(cfg=sim_l, resources=res, products=prod, policy="dbr", seed=806309, patches=["P2", "P3"])
    # resource `queue` records (log_queues=True; the shipped configs misspell the key)
    simq, resq, prodq, _ = _example("simple_scheduler")
    simq.pop("print_mode", None)
    simq.update(run_until=300, warmup=40, memory_size=100000, log_queues=True)
    out["ss_queues"] = dict(cfg=simq, resources=resq, products=prodq, policy="base", seed=806309, patches=[])
    # --- combinations no shipped config exercises (added late in round 1)
    # max-priority dispatch x onDue delivery (patch P1): per-order delivery timers next to the priority scan
    simo, reso, prodo, _ = _example("simple_scheduler")
    simo.pop("print_mode", None)
    simo.update(run_until=450, warmup=60, memory_size=100000, delivery_mode="onDue")
    out["ss_ondue"] = dict(cfg=simo, resources=reso, products=prodo, policy="simple_scheduler", seed=427023,
                           patches=["P1"])
    # 20-resource job shop: lots > 1 (batch sums), queue records, frequent breakdowns, `instantly` delivery
    cfg, res20c, prod20c = sc.jobshop20(run_until=700, warmup=100, mode="instantly")
    cfg["log_queues"] = True
    for i, r in enumerate(res20c.values()):
        r["tbf"] = {"dist": "expo", "params": [90 + 11 * i]}
        r["ttr"] = {"dist": "gamma", "params": [4, 3]}
    for j, p_ in enumerate(prod20c.values()):
        p_["demand"]["quantity"] = {"dist": "uniform", "params": [2, 0.5]}
        p_["demand"]["freq"] = {"dist": "erlang", "params": [24 + j, 12]}
    out["js20_lots_queues"] = dict(cfg=cfg, resources=res20c, products=prod20c, policy="base", seed=777001, patches=[])
    # drum-buffer-rope x onDue delivery (P1 + P2 + P3): rope delay timers and delivery timers in one calendar
    simd, resd, prodd, _ = _example("toc_dbr")
    simd.pop("print_mode", None)
    simd.update(run_until=1501, warmup=300, memory_size=200000, delivery_mode="onDue")
    out["dbr_ondue"] = dict(cfg=simd, resources=resd, products=prodd, policy="dbr", seed=280679,
                            patches=["P1", "P2", "P3"])
    # python-typed clock vs float32 Timeouts that round back onto the current timestamp: positive setups far below the
    # float32 spacing of the clock (1e-4 against 4.9e-4 beyond t = 4096) on every other resource, zero setups elsewhere;
    # orders released at int-typed tick instants then meet `now + np.float32(1e-4) == now` with a clock that turns
    # float32 (found at t = 128064 of the shipped horizon; the kernel kept the python type there before the fix)
    simy, resy, prody, _ = _example("toc_dbr")
    simy.pop("print_mode", None)
    simy.update(run_until=7001, warmup=4200, memory_size=200000, scheduler_interval=24, cb_target_level=120)
    for i, r in enumerate(resy.values()):
        r["setup"] = {"dist": "constant", "params": [1e-4 if i % 2 == 0 else 0, 0]}
    out["dbr_tiny_setup"] = dict(cfg=simy, resources=resy, products=prody, policy="dbr", seed=806309,
                                 patches=["P2", "P3"])
    # calendar entries of different clock types tied at one float32 instant: the events each of them schedules keep
    # their creator's type (Environment.step() takes _now from the popped entry); same stress configuration as
    # dbr_typed, a seed and horizon where the kernel of the round-1 GPU runs logged a float64-path utilisation at
    # t = 2184.97 (42 differing records up to t = 2301)
    out["dbr_mixed_types"] = dict(cfg=dict(sim_t, run_until=2401, warmup=100), resources=res_t, products=prod,
                                  policy="dbr", seed=4242, patches=["P2", "P3"])
    # several python-typed (float64) instants inside ONE float32 spacing of the clock: the reference's heap compares two
    # python floats exactly, so the earlier float64 instant finishes its whole zero-delay cascade before the later one
    # starts, whatever their eids (a kernel that orders the calendar by float32 time alone interleaves them by eid:
    # with this seed it diverges at t = 4322.37 -- 39 844 differing records and another step count up to t = 8001)
    out["dbr_f64_order"] = dict(cfg=dict(sim_t, run_until=5001, warmup=4000), resources=res_t, products=prod,
                                policy="dbr", seed=8, patches=["P2", "P3"])
    # ORACLE-ONLY goldens (tests/conftest.py ORACLE_ONLY_NAMES): min_lotsize != 0 (toc_dbr/simulation.py:222-229) turns the
    # lot into a python int / float, the CCR time into a float64 product and the rope delay into a python-typed Timeout.
    # The kernel refuses the option; the traveling oracle is pinned on it all the same.
    out["dbr_min_lot"] = dict(cfg=dict(sim, run_until=3001, warmup=300, min_lotsize=2), resources=res, products=prod,
                              policy="dbr", seed=1234, patches=["P2", "P3"])
    out["dbr_min_lot_frac"] = dict(cfg=dict(sim, run_until=2001, warmup=300, min_lotsize=2.5), resources=res,
                                   products=prod, policy="dbr", seed=99, patches=["P2", "P3"])
    # zero setups and zero-time operations under max-priority dispatch: the machine finishes whole operations inside one
    # timestamp while the transport chain of the previous order is still being processed (event order within a
    # timestamp; the kernel's "machine chain ahead of the transport chain" shortcut needed OP_HOP for this)
    simz, resz, prodz, _ = _example("simple_scheduler")
    simz.pop("print_mode", None)
    simz.update(run_until=400, warmup=40, memory_size=100000)
    for r in resz.values():
        r["setup"] = {"dist": "constant", "params": [0, 0]}
    for j, p_ in enumerate(prodz.values()):
        for k, st in enumerate(p_["processes"].values()):
            if (j + k) % 3 == 0:
                st["processing_time"] = {"dist": "constant", "params": [0, 0]}
    out["ss_zero_ops"] = dict(cfg=simz, resources=resz, products=prodz, policy="simple_scheduler", seed=91421, patches=[])
    # SimPy ValueError cases: negative batch delay, Container.put(0)
    cfg, res, prod = sc.toy2(run_until=400)
    prod["p1"]["processes"]["s1"]["processing_time"] = {"dist": "uniform", "params": [0.2, 2.0]}   # a < 0
    out["err_negdelay"] = dict(cfg=cfg, resources=res, products=prod, policy="base", seed=5, patches=[])
    cfg, res, prod = sc.toy2(run_until=400)
    prod["p1"]["demand"]["quantity"] = {"dist": "uniform", "params": [0.9, 0.5]}                   # round(x,0) hits 0
    out["err_amount"] = dict(cfg=cfg, resources=res, products=prod, policy="base", seed=5, patches=[])
    return out


def write_examples():
    for name in ("simple_scheduler", "toc_dbr"):
        sim, res, prod, exp = rr.load_example(name)
        blob = {"simulation": sim, "resources": res, "products": prod, "experiment": exp,
                "source": f"/root/reference/examples/{name} (config.yaml + products + resources), dumped by oracle/gen_golden.py"}
        (GOLDEN / f"example_{name}.json").write_text(json.dumps(blob, separators=(",", ":")))


def write_variate_kat():
    """Known-answer vectors from the reference's own DistributionGenerator (engine/utils.py:21-105)."""
    DG = rr.load_reference()["utils"].DistributionGenerator
    cases = [("erlang", [10, 5.77, 3]), ("uniform", [10, 5.77, 0]), ("expo", [4635]), ("normal", [5.24, 2.97]),
             ("gamma", [0.78, 0.45]), ("constant", [63]), ("normal", [0.04, 0.05]), ("constant", [0, 0])]
    batch = [("gamma", [0.78, 0.45], 1), ("gamma", [0.78, 0.45], 5), ("expo", [2.0], 3), ("normal", [1.0, 2.0], 4),
             ("uniform", [10, 5.77], 2), ("constant", [0.78], 3), ("erlang", [1.2, 0.7], 0)]
    kat = {"scalar": [], "batch": [], "run_seeds": {}}
    for seed in (42, 54907):
        g = DG(seed)
        for _ in range(3):
            for dist, params in cases:
                v = g.random_number(dist, params)
                kat["scalar"].append({"seed": seed, "dist": dist, "params": params, "value": float(v),
                                      "is_int": isinstance(v, int)})
        g = DG(seed)
        for _ in range(2):
            for dist, params, n in batch:
                kat["batch"].append({"seed": seed, "dist": dist, "params": params, "count": n,
                                     "value": g.sum_random_numbers_batch(dist, params, n)})
    for seed in (123, 20260101):
        g = DG(seed)
        kat["run_seeds"][str(seed)] = [g.random_int() for _ in range(8)]
    (GOLDEN / "variates_kat.json").write_text(json.dumps(kat, indent=0))


def generate(name, spec):
    out = rr.run_reference(spec["cfg"], spec["resources"], spec["products"], spec["seed"],
                           policy=spec["policy"], patches=spec["patches"])
    scen = {k: spec[k] for k in ("cfg", "resources", "products", "policy", "seed", "patches")}
    sim = out["sim"]
    metrics = {"products": sim.products_metrics().to_dict(orient="list"),
               "resources": sim.resources_metrics().to_dict(orient="list"),
               "general": sim.general_metrics().to_dict(orient="list")}
    np.savez_compressed(
        GOLDEN / f"{name}.npz",
        ring=out["ring"], time=out["time"], value=out["value"],
        tape_A=out["tape_A"], tape_B=out["tape_B"], tape_C=out["tape_C"], tape_D=out["tape_D"],
        steps=np.int64(out["steps"]), error=np.str_(out["error"] or ""),
        scenario=np.str_(json.dumps(scen)), metrics=np.str_(json.dumps(metrics)),
        now=np.float64(out["now"]), now_is_int=np.bool_(isinstance(out["now"], int)))
    print(f"{name}: steps={out['steps']} records={len(out['ring'])} tapes="
          f"{len(out['tape_A'])}/{len(out['tape_B'])}/{len(out['tape_C'])}/{len(out['tape_D'])} "
          f"error={out['error']}")


def main(argv):
    GOLDEN.mkdir(parents=True, exist_ok=True)
    write_examples()
    write_variate_kat()
    table = scenario_table()
    for name in (argv or table.keys()):
        generate(name, table[name])


if __name__ == "__main__":
    main(sys.argv[1:])
