"""CPU: host-side drop-in surface -- Logger view semantics (ring wrap quirk E2), hook policing, run-seed derivation,
shard bounds, moments -> statistics."""
import numpy as np
import pandas as pd
import pytest

from manusim_b200 import compile_scenario, scenarios
from manusim_b200.experiment import generate_run_seeds, shard_bounds, stats_from_moments
from manusim_b200.logs import Logger
from manusim_b200.metrics import (MetricGeneral, MetricProducts, MetricResources, complete_wide_metrics_df,
                                  frames_from_values, metric_values)


def _tables():
    cfg, res, prod = scenarios.toy2()
    return compile_scenario(cfg, res, prod)


@pytest.mark.parametrize("mem,with_path", [(1000, False), (7, False), (7, True), (1, False)])
def test_load_stream_equals_sequential_log_calls(tmp_path, mem, with_path):
    """Folding the stream must leave exactly the state successive Logger.log() calls leave (logs.py:44-66)."""
    t = _tables()
    names = t.ring_names()
    rng = np.random.default_rng(3)
    n = 200
    ring = rng.integers(0, len(names), size=n).astype(np.uint32)
    tm = np.sort(rng.random(n).astype(np.float32) * 100)
    val = rng.random(n).astype(np.float32)
    a = Logger.for_tables(t, tmp_path / "a" if with_path else None, mem)
    b = Logger.for_tables(t, tmp_path / "b" if with_path else None, mem)
    a.load_stream(names, ring, tm, val)
    for r, x, y in zip(ring, tm, val):
        b.log(names[r][0], names[r][1], (x, y))
    for variable, key in names:
        assert a.log_index[variable][key] == b.log_index[variable][key]
        assert np.array_equal(getattr(a, variable)[key], getattr(b, variable)[key]), (variable, key)
        if with_path:
            fa, fb = tmp_path / "a" / f"log_{variable}_{key}.csv", tmp_path / "b" / f"log_{variable}_{key}.csv"
            assert fa.exists() == fb.exists()
            if fa.exists():
                pd.testing.assert_frame_equal(pd.read_csv(fa), pd.read_csv(fb))
    assert a.get_last_log_value(*names[int(ring[-1])]) == b.get_last_log_value(*names[int(ring[-1])])


def test_run_seeds_match_reference():
    assert generate_run_seeds(123, 5) == [54907, 280679, 91421, 806309, 427023]


def test_shard_bounds_partition():
    for n in (1, 7, 1000, 8_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_stats_from_moments():
    x = np.array([1.0, 2.0, 4.0])
    mom = np.array([[x.sum(), (x * x).sum(), 3.0], [0.0, 0.0, 0.0]])
    st = stats_from_moments(mom)
    assert st["mean"][0] == pytest.approx(x.mean()) and st["std"][0] == pytest.approx(x.std(ddof=1))
    assert np.isnan(st["mean"][1])


def test_metric_values_and_frames():
    t = _tables()
    K = t.n_rings
    s = np.arange(K, dtype=np.float64) + 1
    c = np.ones(K, dtype=np.int64) * 2
    c[3] = 0
    v = metric_values(s, c, t, span=10.0)
    P, R = t.P, t.R
    assert v[0] == 1.0                         # deliveredOntime: sum
    assert np.isnan(v[3])                      # flowTime: mean of nothing
    assert v[11 * P] == s[11 * P] / 10.0       # utilization: sum / span
    assert v[4] == s[4] / 2
    prod, res, gen = frames_from_values(v, t)
    assert list(prod.columns) == ["key"] + [m.name for m in MetricProducts]
    assert list(res.columns) == ["key"] + [m.name for m in MetricResources]
    assert list(gen.columns) == ["key"] + [m.name for m in MetricGeneral]


def test_complete_wide_fill_rules():
    df = pd.DataFrame({"utilization": [0.5]}, index=pd.Index(["rA"], name="key"))
    out = complete_wide_metrics_df(df, list(MetricResources), ["rA", "rB"])
    assert out.loc[1, "utilization"] == 0.0 and out.loc[1, "breakdown"] == 0.0 and np.isnan(out.loc[1, "queue"])


def test_unknown_hook_override_is_refused():
    from manusim_b200.factory_sim import DBRSimulation, FactorySimulation, NewSimulation

    cfg, res, prod = scenarios.toy2()

    class Custom(FactorySimulation):
        def order_selection(self, resource):
            yield None

    with pytest.raises(NotImplementedError, match="recognised policies"):
        Custom(cfg, res, prod)

    class Custom2(NewSimulation):
        def _custom_fg_reduced(self, product):
            print("x")

    with pytest.raises(NotImplementedError):
        Custom2(cfg, res, prod)
    sim = NewSimulation(cfg, res, prod)                # recognised classes construct on CPU (tables only)
    assert sim.policy == "max_priority" and sim.logs.log_index["wip"]["p1"] == 0
    sim2, r2, p2, _ = scenarios.load_example("toc_dbr")
    d = DBRSimulation(sim2, r2, p2)
    assert d.contraint_resource == "r09" and d.scheduler_interval == 48
    with pytest.raises(ValueError, match="run_until must be specified"):
        FactorySimulation({}, res, prod)


def test_sweep_grid_and_flat_slices():
    from manusim_b200.sweep import expand_grid, flat_slices

    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    base = {"simulation": sim, "resources": res, "products": prod}
    pts = expand_grid(base, {"products.*.demand.freq.params.0": [12.5, 10.0],
                             "products.*.shipping_buffer": [50, 100, 150]})
    assert len(pts) == 6
    assert pts[0][0] == {"products.*.demand.freq.params.0": 12.5, "products.*.shipping_buffer": 50}
    assert all(p["shipping_buffer"] == 150 for p in pts[2][1]["products"].values())
    assert all(p["demand"]["freq"]["params"][0] == 10.0 for p in pts[5][1]["products"].values())
    assert base["products"]["produto01"]["shipping_buffer"] == 200          # base tree untouched
    # the flattened (scenario, replication) space is covered exactly once for any world size
    for n_s, reps, world in ((6, 100, 1), (6, 100, 4), (256, 31250, 8), (3, 7, 5)):
        seen = np.zeros((n_s, reps), dtype=int)
        for rank in range(world):
            for s, first, cnt in flat_slices(n_s, reps, world, rank):
                seen[s, first:first + cnt] += 1
        assert (seen == 1).all()


def test_sweep_kernel_family_groups_scenarios_by_the_kernel_they_launch():
    """SweepRunner launches kernel family by kernel family: scenarios that differ only in numbers share a family, a
    different delivery mode, policy, queue-record switch or an off-standard shape does not."""
    from types import SimpleNamespace as NS

    from manusim_b200 import compile_scenario
    from manusim_b200.sweep import kernel_family

    def fam(cfg, res, prod, const_layout=1, **kw):
        return kernel_family(NS(tables=compile_scenario(cfg, res, prod, **kw), engine=NS(info={"const_layout": const_layout})))

    cfg, res, prod = scenarios.jobshop20(run_until=300, warmup=30, mode="asReady")
    a = fam(cfg, res, prod)
    slower = {k: dict(v, demand=dict(v["demand"], freq={"dist": "expo", "params": [14.0]})) for k, v in prod.items()}
    assert fam(cfg, res, slower) == a
    assert fam(dict(cfg, delivery_mode="onDue"), res, prod) != a
    assert fam(cfg, res, prod, policy="max_priority") != a
    assert fam(dict(cfg, log_queues=True), res, prod) != a
    assert fam(cfg, res, prod, const_layout=0) != a
    order = sorted([fam(dict(cfg, delivery_mode=m), res, prod) for m in ("onDue", "asReady", "onDue", "asReady")])
    assert order[0] == order[1] and order[2] == order[3] and order[1] != order[2]


def test_monitor_step_count_matches_the_reference_monitor_process():
    """print_mode "metrics" (the constructor default) starts the reference's console monitor: it only prints, but its
    Initialize and Timeout events are Environment.step() calls.  FactorySimulation.monitor_steps() must equal the
    difference the unmodified reference shows between a run with and without the monitor (factory_sim.py:656-684)."""
    import contextlib
    import io

    from manusim_b200.factory_sim import FactorySimulation
    from oracle import run_reference as rr

    if not rr.reference_available():
        pytest.skip("/root/reference not mounted")
    cfg, res, prod = scenarios.toy2(run_until=400, warmup=30)
    ref = rr.load_reference()
    for extra in ({}, {"monitor_warmup": 100, "monitor_interval": 100}, {"monitor_interval": 37}, {"monitor_warmup": 400},
                  {"monitor_warmup": 25, "monitor_interval": 125}):
        c = dict(cfg, **extra)
        steps = {}
        for mode in ("none", "metrics"):
            sim = ref["FactorySimulation"](dict(c), res, prod, print_mode=mode, seed=5)
            sim.reset_simulation(5, None)
            with contextlib.redirect_stdout(io.StringIO()):
                sim.run_simulation()
            steps[mode] = next(sim.env._eid) - len(sim.env._queue)
        ours = FactorySimulation.__new__(FactorySimulation)
        ours.print_mode, ours.run_until = "metrics", c["run_until"]
        ours.monitor_warmup = c.get("monitor_warmup", c.get("warmup", 0))
        ours.monitor_interval = c.get("monitor_interval", int(c["run_until"] / 4))
        assert ours.monitor_steps() == steps["metrics"] - steps["none"], (extra, steps)
        ours.print_mode = "none"
        assert ours.monitor_steps() == 0


def test_sb_update_and_cb_update_change_nothing_but_the_step_count_in_the_reference():
    """examples/toc_dbr sb_update / cb_update as shipped (see manusim_b200/compiler.py): the unmodified reference produces
    the same records and variates with and without them; sb_update adds the sleeping adjust_shipping_buffer processes'
    steps, which DBRSimulation.host_only_steps() must reproduce."""
    from conftest import load_golden
    from manusim_b200.factory_sim import DBRSimulation
    from oracle import run_reference as rr

    if not rr.reference_available():
        pytest.skip("/root/reference not mounted")
    g = load_golden("dbr_repaired")
    s = g["scenario"]
    for horizon in ({}, {"run_until": 2501, "warmup": 1541}):
        base_cfg = dict(s["cfg"], **horizon)
        base = rr.run_reference(base_cfg, s["resources"], s["products"], s["seed"], policy="dbr", patches=s["patches"])
        for extra in ({"sb_update": True}, {"cb_update": True}, {"sb_update": True, "cb_update": True}):
            cfg = dict(base_cfg, **extra)
            o = rr.run_reference(cfg, s["resources"], s["products"], s["seed"], policy="dbr", patches=s["patches"])
            for k in ("ring", "time", "value", "tape_A", "tape_B", "tape_C", "tape_D"):
                assert np.array_equal(o[k], base[k]), (extra, k)
            ours = DBRSimulation(cfg, s["resources"], s["products"], print_mode="none", dbr_repair=True)
            plain = DBRSimulation(base_cfg, s["resources"], s["products"], print_mode="none", dbr_repair=True)
            assert ours.host_only_steps() - plain.host_only_steps() == o["steps"] - base["steps"], (extra, horizon)
