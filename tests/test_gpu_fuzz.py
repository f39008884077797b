"""GPU: randomised scenarios (routes with repeated / consecutive resources, zero and random setups, lots > 1, zero-time
operations, frequent breakdowns, every delivery mode and dispatch policy).  The kernel runs in Philox mode and records
the variates it drew; the CPU oracle replays them and must reproduce the record stream and the SimPy step count bit
for bit -- or fail with the same SimPy ValueError class."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dist(rng, scale, allow_zero=False):
    kind = rng.choice(["constant", "uniform", "gamma", "erlang", "expo", "normal"])
    m = float(scale * rng.uniform(0.3, 1.5))
    if kind == "constant":
        return {"dist": "constant", "params": [0 if (allow_zero and rng.random() < 0.4) else round(m, 3)]}
    if kind == "uniform":
        return {"dist": "uniform", "params": [round(m, 3), round(m * rng.uniform(0.05, 0.5), 3)]}     # a > 0
    if kind in ("gamma", "erlang"):
        return {"dist": kind, "params": [round(m, 3), round(m * rng.uniform(0.3, 1.6), 3)]}           # shape < 1 too
    if kind == "expo":
        return {"dist": "expo", "params": [round(m, 3)]}
    return {"dist": "normal", "params": [round(m, 3), round(m * rng.uniform(0.1, 1.2), 3)]}          # clamps at 0


def random_scenario(seed):
    rng = np.random.default_rng(seed)
    R, P = int(rng.integers(2, 7)), int(rng.integers(1, 5))
    resources = {}
    for r in range(R):
        res = {"quantity": int(rng.integers(1, 3)), "setup": _dist(rng, 0.3, allow_zero=True)}
        if rng.random() < 0.5:
            res["tbf"] = _dist(rng, 40.0)
            res["ttr"] = _dist(rng, 3.0, allow_zero=True)
        resources[f"m{r}"] = res
    products = {}
    for p in range(P):
        n = int(rng.integers(1, 7))
        procs = {f"s{j}": {"resource": f"m{int(rng.integers(0, R))}", "processing_time": _dist(rng, 1.2, allow_zero=True)}
                 for j in range(n)}
        qty = {"dist": "constant", "params": [int(rng.integers(1, 4))]} if rng.random() < 0.6 else \
            {"dist": "uniform", "params": [3, 0.5]}
        products[f"p{p}"] = {"processes": procs, "shipping_buffer": int(rng.integers(5, 30)),
                             "demand": {"freq": _dist(rng, 6.0 * P), "quantity": qty, "duedate": _dist(rng, 25.0)}}
    mode = str(rng.choice(["asReady", "onDue", "instantly"]))
    policy = str(rng.choice(["fifo", "max_priority", "edd", "dbr"]))
    cfg = {"run_until": int(rng.integers(150, 400)), "warmup": int(rng.choice([0, 20, 50])), "delivery_mode": mode,
           "log_queues": bool(rng.random() < 0.3), "scheduler_interval": int(rng.choice([12, 24, 37])),
           "cb_target_level": float(rng.choice([30, 80, 200]))}
    return cfg, resources, products, policy, bool(rng.random() < 0.5)


@pytest.mark.parametrize("seed", list(range(48)))
def test_random_scenario_matches_oracle(seed):
    import torch

    from manusim_b200 import BatchEngine, compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    cfg, res, prod, policy, repair = random_scenario(seed)
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250,
                                       max_demand_orders=250, fifo_capacity=256))
    n = 6
    r = eng.run(n, seed=1000 + seed, n_log_reps=n, log_cap=1 << 16, record_tapes=1 << 14)
    torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    cnt = r.rec["count"].cpu().numpy()
    assert cnt.max() < (1 << 14)
    for i in range(n):
        tp = {k: r.rec[k][i, : cnt[i, j]].cpu().numpy() for j, k in enumerate("ABCD")}
        o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
        ring, tm, val, _ = decode_log(r.log[i].cpu().numpy(), int(r.log_count[i]))
        if o["error"] is not None:
            want = 1 if "Negative delay" in o["error"] else 2
            assert st[i] == want, (seed, i, st[i], o["error"])
        else:
            assert st[i] == 0, (seed, i, st[i])
            assert int(r.n_events[i]) == o["steps"], (seed, i)
        assert np.array_equal(ring, o["ring"].astype(np.uint32)), (seed, i, policy, cfg)
        assert np.array_equal(tm.view(np.uint32), o["time"].view(np.uint32)), (seed, i, policy, cfg)
        assert np.array_equal(val.view(np.uint32), o["value"].view(np.uint32)), (seed, i, policy, cfg)
