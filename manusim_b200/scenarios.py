"""Scenario builders: plain (config, resources, products) dict triples in the reference's YAML schema
(/root/reference/examples/*/conf/**.yaml) for the BASELINE.json workloads.

No engine code here -- these are inputs, shared by tests, bench.py and the oracle harness.
"""
from __future__ import annotations

import copy
import json
import math
from pathlib import Path

_GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _d(dist, *params):
    return {"dist": dist, "params": list(params)}


def toy2(quantity=2, due=20, warmup=0, run_until=200, mode="asReady"):
    """2-resource / 1-product all-constant toy of SURVEY.md Appendix C."""
    cfg = {"run_until": run_until, "warmup": warmup, "delivery_mode": mode, "memory_size": 100000}
    resources = {"rA": {"quantity": 1, "setup": _d("constant", 0, 0)},
                 "rB": {"quantity": 1, "setup": _d("constant", 0.5, 0)}}
    products = {"p1": {"processes": {"s1": {"resource": "rA", "processing_time": _d("constant", 2.0)},
                                     "s2": {"resource": "rB", "processing_time": _d("constant", 3.0)}},
                       "demand": {"freq": _d("constant", 10), "quantity": _d("constant", quantity),
                                  "duedate": _d("constant", due)},
                       "shipping_buffer": 10}}
    return cfg, resources, products


def jobshop20(run_until=10000, warmup=1000, mode="asReady", breakdowns=True):
    """BASELINE config #4: synthetic 20-resource x 10-product job shop with setups,
    TBF/TTR breakdowns; route of product i (1-based) has 8 + (i mod 5) steps, step j on resource
    (3 i + 7 j) mod 20; gamma processing times targeting rho ~ 0.8."""
    R, P = 20, 10
    routes = []
    visits = [0] * R
    for i in range(1, P + 1):
        steps = [(3 * i + 7 * j) % R for j in range(8 + (i % 5))]
        routes.append(steps)
        for r in steps:
            visits[r] += 1
    resources = {}
    for r in range(R):
        res = {"quantity": 1, "setup": _d("normal", 0.05, 0.02)}
        if breakdowns:
            res["tbf"] = _d("expo", 4000)
            res["ttr"] = _d("normal", 5, 2)
        resources[f"r{r + 1:02d}"] = res
    products = {}
    for i, steps in enumerate(routes, start=1):
        procs = {}
        for j, r in enumerate(steps):
            m = 0.8 * 10 / visits[r]
            procs[f"p{i}-{j + 1:02d}"] = {"resource": f"r{r + 1:02d}",
                                          "processing_time": _d("gamma", m, m / math.sqrt(3))}
        products[f"p{i:02d}"] = {"processes": procs, "shipping_buffer": 100,
                                 "demand": {"freq": _d("erlang", 10, 5.77), "quantity": _d("constant", 1),
                                            "duedate": _d("constant", 63)}}
    cfg = {"run_until": run_until, "warmup": warmup, "delivery_mode": mode, "memory_size": 10 ** 7}
    return cfg, resources, products


def make_deterministic(resources, products):
    """BASELINE config #2: every distribution -> constant(params[0]); TBF/TTR removed."""
    resources = copy.deepcopy(resources)
    products = copy.deepcopy(products)
    for r in resources.values():
        r.pop("tbf", None)
        r.pop("ttr", None)
        r["setup"] = _d("constant", r["setup"].get("params", [0])[0])
    for p in products.values():
        for st in p["processes"].values():
            st["processing_time"] = _d("constant", st["processing_time"]["params"][0])
        for k in ("freq", "quantity", "duedate"):
            p["demand"][k] = _d("constant", p["demand"][k]["params"][0])
    return resources, products


def load_example(name):
    """Shipped example configs, dumped verbatim (as JSON, key order preserved) by oracle/gen_golden.py
    from /root/reference/examples/{simple_scheduler,toc_dbr}; returns (simulation, resources, products,
    experiment)."""
    blob = json.loads((_GOLDEN / f"example_{name}.json").read_text())
    return blob["simulation"], blob["resources"], blob["products"], blob["experiment"]
