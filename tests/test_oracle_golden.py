"""CPU: the traveling oracle (oracle/factory_oracle.py on oracle/minisimpy.py) is pinned, record for record,
against golden vectors produced by executing the UNMODIFIED reference (oracle/gen_golden.py)."""
import numpy as np
import pytest

from conftest import ERROR_NAMES, GOLDEN_NAMES, ORACLE_ONLY_NAMES, ORACLE_POLICY, load_golden
from oracle.factory_oracle import metrics_from_trace, run_oracle

FAST = [n for n in GOLDEN_NAMES if n not in ("js20_ondue", "js20_breakdowns")]


def _run(g, tapes=None):
    s = g["scenario"]
    return run_oracle(s["cfg"], s["resources"], s["products"], None if tapes else s["seed"],
                      ORACLE_POLICY[s["policy"]], tapes=tapes, dbr_repair="P3" in s["patches"])


@pytest.mark.parametrize("name", GOLDEN_NAMES + ORACLE_ONLY_NAMES)
def test_oracle_matches_reference_seeded(name):
    g = load_golden(name)
    o = _run(g)
    assert o["error"] is None and g["error"] == ""
    assert o["steps"] == g["steps"]
    for k in ("ring", "time", "value", "tape_A", "tape_B", "tape_C", "tape_D"):
        assert o[k].dtype == g[k].dtype
        assert np.array_equal(o[k].view(np.uint8), g[k].view(np.uint8)), k


@pytest.mark.parametrize("name", FAST + ORACLE_ONLY_NAMES)
def test_oracle_matches_reference_in_replay(name):
    g = load_golden(name)
    o = _run(g, tapes={k: g["tape_" + k] for k in "ABCD"})
    assert o["steps"] == g["steps"]
    for k in ("ring", "time", "value"):
        assert np.array_equal(o[k].view(np.uint8), g[k].view(np.uint8)), k


@pytest.mark.parametrize("name", ["ss_cfg1", "js20_asready", "dbr_repaired"])
def test_metrics_from_trace_match_reference_dataframes(name):
    """per-ring reductions reproduce FactorySimulation.products/resources/general_metrics (factory_sim.py:769-789)."""
    g = load_golden(name)
    s = g["scenario"]
    P, R = len(s["products"]), len(s["resources"])
    _, _, val = metrics_from_trace(g, P, R, s["cfg"]["run_until"], s["cfg"].get("warmup", 0))
    from manusim_b200.compiler import GENERAL_METRICS, PRODUCT_METRICS, RESOURCE_METRICS

    m = g["metrics"]
    for i, name_ in enumerate(PRODUCT_METRICS):
        np.testing.assert_allclose(val[i * P:(i + 1) * P], np.array(m["products"][name_], dtype=float), rtol=2e-6,
                                   equal_nan=True)
    for i, name_ in enumerate(RESOURCE_METRICS):
        np.testing.assert_allclose(val[11 * P + i * R:11 * P + (i + 1) * R], np.array(m["resources"][name_], dtype=float),
                                   rtol=2e-6, equal_nan=True)
    for i, name_ in enumerate(GENERAL_METRICS):
        np.testing.assert_allclose(val[11 * P + 4 * R + i], float(m["general"][name_][0]), rtol=2e-6, equal_nan=True)


def test_appendix_c_microtrace():
    """First records of the 2-resource toy."""
    g = load_golden("toy2")
    P, R = 1, 2
    util_rA = 11 * P + 0 * R + 0
    setup_rB = 11 * P + 2 * R + 1
    flow = 3 * P
    sel = lambda ring: [(float(t), float(v)) for r, t, v in zip(g["ring"], g["time"], g["value"]) if r == ring]
    assert sel(util_rA)[:2] == [(14.0, 4.0), (24.0, 4.0)]
    assert sel(setup_rB)[0] == (14.0, 0.5)
    assert sel(flow)[0] == (20.5, 10.5)


@pytest.mark.parametrize("name", ERROR_NAMES)
def test_oracle_reproduces_reference_value_errors(name):
    """SimPy ValueError cases (negative delay, Container amount <= 0): same records before the abort, same message."""
    g = load_golden(name)
    o = _run(g)
    assert g["error"].startswith("ValueError") and o["error"] == g["error"]
    for k in ("ring", "time", "value", "tape_A", "tape_B", "tape_C"):
        assert np.array_equal(o[k].view(np.uint8), g[k].view(np.uint8)), k
