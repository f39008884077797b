"""GPU parity: the CUDA engine (through the C ABI) must reproduce, bit for bit, the record stream, the SimPy step
count and the per-ring reductions of the golden vectors produced by the unmodified reference, when fed the
reference's own variate tapes (replay mode)."""
import numpy as np
import pytest

from conftest import ERROR_NAMES, GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

POLICY = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}
BASE_NAMES = list(GOLDEN_NAMES)


def _engine(scn, **kw):
    from manusim_b200 import BatchEngine, compile_scenario

    t = compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy=POLICY[scn["policy"]],
                         dbr_repair="P3" in scn["patches"], **kw)
    return BatchEngine(t), t


@pytest.mark.parametrize("name", BASE_NAMES)
def test_replay_trace_bit_exact(name):
    import torch

    from manusim_b200.engine import decode_log

    g = load_golden(name)
    eng, t = _engine(g["scenario"])
    n = 5
    tapes = [{k: g["tape_" + k] for k in "ABCD"}] * n
    cap = len(g["ring"]) + 64
    res = eng.run(n, tapes=tapes, n_log_reps=n, log_cap=cap)
    torch.cuda.synchronize()
    status = res.status.cpu().numpy()
    assert (status == 0).all(), status
    assert (res.n_events.cpu().numpy() == g["steps"]).all(), (res.n_events.cpu().numpy(), g["steps"])
    counts = res.log_count.cpu().numpy()
    logs = res.log.cpu().numpy()
    for i in range(n):
        assert counts[i] == len(g["ring"])
        ring, tm, val, seq = decode_log(logs[i], int(counts[i]))
        assert np.array_equal(ring, g["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), g["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), g["value"].view(np.uint32))
        assert np.array_equal(seq, np.arange(len(seq), dtype=np.uint32))
    # fused reductions == plain float64 sums / counts of the golden records
    K = eng.K
    s = np.zeros(K)
    c = np.zeros(K, dtype=np.int64)
    np.add.at(s, g["ring"].astype(np.int64), g["value"].astype(np.float64))
    np.add.at(c, g["ring"].astype(np.int64), 1)
    acc = res.acc_sum.cpu().numpy()
    cnt = res.acc_cnt.cpu().numpy()
    for i in range(n):
        np.testing.assert_allclose(acc[i], s, rtol=1e-12, atol=1e-9)
        assert np.array_equal(cnt[i].astype(np.int64), c)


def _oracle_vs_gpu_philox(cfg, res, prod, policy, n=24, seed=77, dbr_repair=False, check=(0, 7, 23), **kw):
    """Philox mode: the kernel records the variates it drew; the oracle replays them and must reproduce the record
    stream and the SimPy step count bit for bit (covers configurations that have no reference golden)."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=dbr_repair, **kw))
    r = eng.run(n, seed=seed, n_log_reps=n, log_cap=1 << 17, record_tapes=1 << 15)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all(), r.status.cpu().numpy()
    cnt = r.rec["count"].cpu().numpy()
    assert cnt.max() < (1 << 15)
    for i in check:
        tp = {k: r.rec[k][i, : cnt[i, j]].cpu().numpy() for j, k in enumerate("ABCD")}
        o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=dbr_repair)
        assert o["error"] is None
        ring, tm, val, _ = decode_log(r.log[i].cpu().numpy(), int(r.log_count[i]))
        assert int(r.n_events[i]) == o["steps"]
        assert np.array_equal(ring, o["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), o["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), o["value"].view(np.uint32))
    return r


@pytest.mark.parametrize("policy", ["fifo", "max_priority", "edd"])
@pytest.mark.parametrize("mode", ["asReady", "onDue", "instantly"])
def test_philox_runs_match_oracle_replay(policy, mode):
    from manusim_b200 import scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=700, warmup=100, mode=mode)
    for i, r in enumerate(res.values()):
        r["tbf"] = {"dist": "expo", "params": [90 + 5 * i]}
    for j, p in enumerate(prod.values()):
        p["demand"]["duedate"] = {"dist": "uniform", "params": [40 + 3 * j, 10]}
    _oracle_vs_gpu_philox(cfg, res, prod, policy)


@pytest.mark.parametrize("repair", [False, True])
def test_philox_dbr_matches_oracle_replay(repair):
    from manusim_b200 import scenarios

    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    cfg = dict(sim, run_until=1501, warmup=200)
    cfg.pop("print_mode", None)
    _oracle_vs_gpu_philox(cfg, res, prod, "dbr", n=8, check=(0, 7), dbr_repair=repair)


@pytest.mark.parametrize("name,status", list(zip(ERROR_NAMES, (1, 2))))
def test_value_error_cases_become_status_codes(name, status):
    """What makes SimPy raise inside env.run() is a per-replication status; records before the abort still match."""
    import torch

    from manusim_b200.engine import decode_log

    g = load_golden(name)
    eng, t = _engine(g["scenario"])
    res = eng.run(3, tapes=[{k: g["tape_" + k] for k in "ABCD"}] * 3, n_log_reps=3, log_cap=4096)
    torch.cuda.synchronize()
    assert (res.status.cpu().numpy() == status).all()
    ring, tm, val, _ = decode_log(res.log[1].cpu().numpy(), int(res.log_count[1]))
    assert np.array_equal(ring, g["ring"].astype(np.uint32))
    assert np.array_equal(tm, g["time"]) and np.array_equal(val, g["value"])


def test_factory_simulation_raises_like_simpy():
    from manusim_b200.factory_sim import FactorySimulation

    g = load_golden("err_amount")
    s = g["scenario"]
    sim = FactorySimulation(s["cfg"], s["resources"], s["products"], print_mode="none")
    sim.reset_simulation(s["seed"], None, tapes={k: g["tape_" + k] for k in "ABCD"})
    with pytest.raises(ValueError, match="amount"):
        sim.run_simulation()


def test_config2_ten_thousand_deterministic_replications_equal_the_oracle_trace():
    """BASELINE config #2: every one of 10 000 replications of the all-constant job shop must equal the single
    reference trace (seeds are irrelevant when nothing is random)."""
    import torch

    from manusim_b200.engine import decode_log

    g = load_golden("ss_det")
    eng, t = _engine(g["scenario"])
    n = 10000
    res = eng.run(n, seed=123456, n_log_reps=4, log_cap=len(g["ring"]) + 8)
    torch.cuda.synchronize()
    assert (res.status == 0).all() and (res.n_events == g["steps"]).all()
    assert (res.acc_sum == res.acc_sum[0]).all() and (res.acc_cnt == res.acc_cnt[0]).all()
    for i in range(4):
        ring, tm, val, _ = decode_log(res.log[i].cpu().numpy(), int(res.log_count[i]))
        assert np.array_equal(ring, g["ring"].astype(np.uint32))
        assert np.array_equal(tm, g["time"]) and np.array_equal(val, g["value"])


@pytest.mark.parametrize("workload,run_until", [("jobshop20", 2500), ("simple_det", 1000), ("toc_dbr", 4001)])
def test_benchmarked_fast_kernels_equal_the_oracle_checked_general_kernels(workload, run_until):
    """The kernels bench.py launches (no queue records, no tape recording, standard shape: compile-time slab / table
    offsets, `info["const_layout"] == 1`) against the general kernels that the oracle-replay tests above check (tape
    recording on, run-time offsets): same Philox keys, same pools => identical record streams, step counts and accumulators,
    record for record.  Closes the chain oracle == general kernel == benchmarked kernel on the bench's own workloads."""
    import torch

    import bench
    from manusim_b200 import BatchEngine, compile_scenario

    spec = bench.workload_spec(workload, run_until)
    for i, ((cfg, res, prod), policy, _, _) in enumerate(spec["parts"]):
        mpo, mdo = spec.get("pools", [(0, 0)] * len(spec["parts"]))[i]
        eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False),
                                           max_production_orders=mpo, max_demand_orders=mdo))
        assert eng.info["const_layout"] == 1, "the bench workloads must run on the fast path"
        n, cap = 64, 1 << 17
        fast = eng.run(n, seed=4242 + i, n_log_reps=n, log_cap=cap)
        general = eng.run(n, seed=4242 + i, n_log_reps=n, log_cap=cap, record_tapes=1 << 16)
        torch.cuda.synchronize()
        st = fast.status.cpu().numpy()
        assert np.array_equal(st, general.status.cpu().numpy())
        ok = st == 0                      # a pool overflow stops both kernels at the same event
        assert ok.sum() >= n - 4, st
        assert np.array_equal(fast.n_events.cpu().numpy(), general.n_events.cpu().numpy())
        lc = fast.log_count.cpu().numpy()
        assert np.array_equal(lc, general.log_count.cpu().numpy()) and lc.max() <= cap
        fl, gl = fast.log.cpu().numpy(), general.log.cpu().numpy()
        for r in np.nonzero(ok)[0]:
            assert np.array_equal(fl[r, : lc[r]], gl[r, : lc[r]]), (workload, i, int(r))
        assert np.array_equal(fast.acc_cnt.cpu().numpy()[ok], general.acc_cnt.cpu().numpy()[ok])
        assert np.array_equal(fast.acc_sum.cpu().numpy()[ok].view(np.uint64), general.acc_sum.cpu().numpy()[ok].view(np.uint64))
        eng.close()
