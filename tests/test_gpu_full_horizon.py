"""GPU parity at BASELINE.json's FULL horizons, on the compiled SASS (round 1 checked these horizons only on the CPU-fiber
emulation of the kernel source): config #4 both delivery parts at run_until = 10 000, config #3 (examples/toc_dbr,
repaired and as shipped) at run_until = 200 001, config #2 (deterministic simple_scheduler) at run_until = 1 000 with
10 000 replications.  Philox runs record the variates they drew; the oracle replays them and must reproduce the complete
record stream (ring, float32 time, float32 value, emission order) and the SimPy step count bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_against_oracle_replay(cfg, res, prod, policy, repair, n, seed, cap_log=1 << 21, cap_tape=1 << 21):
    import torch

    from manusim_b200 import BatchEngine, compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250,
                                       max_demand_orders=250, fifo_capacity=256))
    r = eng.run(n, seed=seed, n_log_reps=n, log_cap=cap_log, record_tapes=cap_tape)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all(), r.status.cpu().numpy()
    cnt = r.rec["count"].cpu().numpy()
    lc = r.log_count.cpu().numpy()
    assert cnt.max() < cap_tape and lc.max() <= cap_log
    out = []
    for i in range(n):
        tp = {k: r.rec[k][i, : cnt[i, j]].cpu().numpy() for j, k in enumerate("ABCD")}
        o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
        assert o["error"] is None
        ring, tm, val, seq = decode_log(r.log[i].cpu().numpy(), int(lc[i]))
        assert int(r.n_events[i]) == o["steps"], (i, int(r.n_events[i]), o["steps"])
        assert len(ring) == len(o["ring"])
        assert np.array_equal(ring, o["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), o["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), o["value"].view(np.uint32))
        assert np.array_equal(seq, np.arange(len(seq), dtype=np.uint32))
        out.append((o["steps"], len(ring)))
    eng.close()
    return out


@pytest.mark.parametrize("part", [0, 1])
def test_config4_full_horizon_both_delivery_parts(part):
    """BASELINE config #4 (20-resource job shop, run_until 10 000): asReady (part 0) and onDue (part 1)."""
    import bench

    spec = bench.workload_spec("jobshop20", None)
    (cfg, res, prod), policy, _, _ = spec["parts"][part]
    assert cfg["run_until"] == 10000
    out = _check_against_oracle_replay(cfg, res, prod, policy, False, n=2, seed=20261020 + part)
    assert all(steps > 1_200_000 for steps, _ in out)


@pytest.mark.parametrize("workload", ["toc_dbr", "toc_dbr_shipped"])
def test_config3_full_horizon_repaired_and_as_shipped(workload):
    """BASELINE config #3 (examples/toc_dbr, run_until 200 001, warm-up 100 000): python-typed clock included."""
    import bench

    spec = bench.workload_spec(workload, None)
    (cfg, res, prod), policy, _, _ = spec["parts"][0]
    assert cfg["run_until"] == 200001
    out = _check_against_oracle_replay(cfg, res, prod, policy, spec["dbr_repair"], n=1, seed=31337,
                                       cap_log=1 << 22, cap_tape=1 << 22)
    assert out[0][0] > 1_000_000


def test_config2_full_horizon_ten_thousand_deterministic_replications():
    """BASELINE config #2 at its stated horizon (run_until 1 000): 10 000 replications of the all-constant
    simple_scheduler model, every one equal to the oracle's single trace -- 122 421 SimPy steps."""
    import torch

    import bench
    from manusim_b200 import BatchEngine, compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    spec = bench.workload_spec("simple_det", None)
    (cfg, res, prod), policy, _, _ = spec["parts"][0]
    assert cfg["run_until"] == 1000
    o = run_oracle(cfg, res, prod, 1, policy)
    assert o["error"] is None and o["steps"] == 122421
    eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy))
    n = 10000
    r = eng.run(n, seed=987654321, n_log_reps=3, log_cap=len(o["ring"]) + 8)
    torch.cuda.synchronize()
    assert (r.status == 0).all() and (r.n_events == o["steps"]).all()
    assert (r.acc_sum == r.acc_sum[0]).all() and (r.acc_cnt == r.acc_cnt[0]).all()
    for i in range(3):
        ring, tm, val, _ = decode_log(r.log[i].cpu().numpy(), int(r.log_count[i]))
        assert np.array_equal(ring, o["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), o["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), o["value"].view(np.uint32))
