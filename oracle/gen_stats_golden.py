"""Generate tests/golden/stats_*.npz: per-ring statistics of N replications of the UNMODIFIED reference (its own MT19937 /
PCG64 generators), the CPU side of the Philox-mode statistical parity tests (BASELINE.json north_star: "per-metric means
over N replications must agree with the reference within a stated confidence-interval tolerance"; SURVEY 8(d): >= 200
CPU replications, shortened horizon for config #3).

TEST INFRASTRUCTURE, build container only (needs /root/reference):
    python -m oracle.gen_stats_golden [name ...] [--reps 256] [--workers 8]

Each file holds, per ring (variable, key): the per-replication metric values' mean, sample variance and count of
non-NaN values over the replications (NaN = mean-type metric without a record in that replication), the number of
replications, the mean SimPy step count, and the JSON scenario (config / resources / products / policy / patches /
experiment seed) so that the GPU test rebuilds the same scenario on a box without /root/reference.
"""
from __future__ import annotations

import argparse
import copy
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def scenarios_table():
    from manusim_b200 import scenarios as sc
    from oracle import run_reference as rr

    out = {}
    sim, res, prod, _ = rr.load_example("simple_scheduler")
    sim = dict(copy.deepcopy(sim), memory_size=20000)
    sim.pop("print_mode", None)
    # config #1 as shipped (run_until 1000, warm-up 50, max-priority dispatch)
    out["stats_ss_cfg1"] = dict(cfg=sim, resources=res, products=prod, policy="simple_scheduler", patches=[], exp_seed=123)
    # config #3 shortened: examples/toc_dbr repaired, run_until 20 001 / warm-up 2 000
    simd, resd, prodd, _ = rr.load_example("toc_dbr")
    simd = dict(copy.deepcopy(simd), run_until=20001, warmup=2000, memory_size=20000)
    simd.pop("print_mode", None)
    out["stats_dbr_cfg3_short"] = dict(cfg=simd, resources=resd, products=prodd, policy="dbr", patches=["P2", "P3"],
                                       exp_seed=4242)
    # config #4 shortened: 20-resource job shop with breakdowns, both delivery modes, run_until 2 500 / warm-up 250,
    # breakdowns frequent enough to be seen inside the shortened horizon
    for mode in ("asReady", "onDue"):
        cfg, r20, p20 = sc.jobshop20(run_until=2500, warmup=250, mode=mode)
        cfg["memory_size"] = 20000
        for i, r in enumerate(r20.values()):
            r["tbf"] = {"dist": "expo", "params": [600 + 20 * i]}
        out[f"stats_js20_{mode.lower()}"] = dict(cfg=cfg, resources=r20, products=p20, policy="base",
                                                 patches=["P1"] if mode == "onDue" else [], exp_seed=20260101)
    return out


def _one(task):
    name, spec, seed = task
    from oracle import run_reference as rr
    from oracle.factory_oracle import metrics_from_trace

    out = rr.run_reference(spec["cfg"], spec["resources"], spec["products"], seed, policy=spec["policy"],
                           patches=spec["patches"])
    assert out["error"] is None, out["error"]
    P, R = len(spec["products"]), len(spec["resources"])
    _, _, val = metrics_from_trace(out, P, R, spec["cfg"]["run_until"], spec["cfg"].get("warmup", 0))
    return val, out["steps"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--workers", type=int, default=8)
    a = ap.parse_args()
    from concurrent.futures import ProcessPoolExecutor

    from oracle.variates import run_seeds

    table = scenarios_table()
    for name in (a.names or table.keys()):
        spec = table[name]
        seeds = [int(s) for s in run_seeds(spec["exp_seed"], a.reps)]
        with ProcessPoolExecutor(max_workers=a.workers) as pool:
            rows = list(pool.map(_one, [(name, spec, s) for s in seeds], chunksize=2))
        vals = np.array([r[0] for r in rows])
        steps = np.array([r[1] for r in rows], dtype=np.float64)
        ok = ~np.isnan(vals)
        n_valid = ok.sum(0)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = np.where(n_valid > 0, np.where(ok, vals, 0).sum(0) / np.maximum(n_valid, 1), np.nan)
            var = np.where(n_valid > 1, (np.where(ok, (vals - mean) ** 2, 0)).sum(0) / np.maximum(n_valid - 1, 1), np.nan)
        np.savez_compressed(GOLDEN / f"{name}.npz", mean=mean, var=var, n_valid=n_valid.astype(np.int64),
                            n_reps=np.int64(a.reps), steps_mean=np.float64(steps.mean()), steps_var=np.float64(steps.var(ddof=1)),
                            scenario=np.str_(json.dumps(spec)),
                            source=np.str_("unmodified /root/reference on oracle/minisimpy.py, seeds = DistributionGenerator"
                                           "(exp_seed).random_int() (experiment.py:138-139); oracle/gen_stats_golden.py"))
        print(f"{name}: {a.reps} replications, mean steps {steps.mean():.0f}, rings with values in every replication "
              f"{int((n_valid == a.reps).sum())}/{len(mean)}", flush=True)


if __name__ == "__main__":
    main()
