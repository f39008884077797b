"""CPU: the kernel SOURCE (manusim_b200/csrc/*.cu, unmodified) compiled by g++ against the SIMT shim in tests/emu and
run on fibers (test infrastructure; the product never loads it).  Three jobs:
  * replay of the reference's tapes through the emulated kernel must give the golden record streams bit for bit --
    event-order logic is checked before any GPU minute is spent, and a divergent collective (a GPU hang) is reported
    instead of hanging a GPU box;
  * Philox-mode runs record their tapes, the oracle replays them (same check as tests/test_gpu_fuzz.py);
  * the same source runs under AddressSanitizer + UBSan (compute-sanitizer is closed on the GPU pool), with a
    negative control proving the sanitizer sees the kernel's stores.
The GPU tests remain the parity gate for the compiled SASS; this file checks the source they are compiled from."""
import json
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import ERROR_NAMES, GOLDEN_NAMES, load_golden

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or os.uname().machine != "x86_64",
                                reason="the SIMT emulation needs g++ on x86-64")

POLICY = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}
FULL = bool(os.environ.get("MSIM_EMU_FULL"))


_asan_build = {}


@pytest.fixture(scope="module")
def emu():
    import threading

    import emu as m

    # the sanitizer flavour compiles while the plain-flavour tests run (different output files)
    def _bg():
        try:
            _asan_build["path"] = m.build("asan")
        except Exception as exc:            # re-raised by the sanitizer test
            _asan_build["error"] = exc

    t = threading.Thread(target=_bg, daemon=True)
    t.start()
    _asan_build["thread"] = t
    m.build("plain")
    return m


# which instantiation a run launches: "auto" = what the library picks by itself (the lean OPTS = 0 kernels when the run
# uses neither queue records nor tape recording), "general" = the OPTS = 3 kernels forced with MSIM_NO_LEAN=1
KERNELS = ["auto", "general"]


def _engine(emu, scn, defines=(), **kw):
    from manusim_b200 import compile_scenario

    t = compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy=POLICY[scn["policy"]],
                         dbr_repair="P3" in scn["patches"], **kw)
    return emu.EmuEngine(t, defines=defines)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_emulated_kernel_replays_reference_goldens_bit_exact(emu, name, kernel, monkeypatch):
    from manusim_b200.engine import decode_log

    monkeypatch.setenv("MSIM_NO_LEAN", "1" if kernel == "general" else "")
    g = load_golden(name)
    eng = _engine(emu, g["scenario"])
    n = 2
    res = eng.run(n, tapes=[{k: g["tape_" + k] for k in "ABCD"}] * n, n_log_reps=n, log_cap=len(g["ring"]) + 64)
    assert (res.status == 0).all(), res.status
    assert (res.n_events == g["steps"]).all(), (res.n_events, g["steps"])
    K = eng.K
    s = np.zeros(K)
    c = np.zeros(K, dtype=np.int64)
    np.add.at(s, g["ring"].astype(np.int64), g["value"].astype(np.float64))
    np.add.at(c, g["ring"].astype(np.int64), 1)
    for i in range(n):
        assert res.log_count[i] == len(g["ring"])
        ring, tm, val, seq = decode_log(res.log[i], int(res.log_count[i]))
        assert np.array_equal(ring, g["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), g["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), g["value"].view(np.uint32))
        assert np.array_equal(seq, np.arange(len(seq), dtype=np.uint32))
        np.testing.assert_allclose(res.acc_sum[i], s, rtol=1e-12, atol=1e-9)
        assert np.array_equal(res.acc_cnt[i].astype(np.int64), c)


def test_float64_calendar_order_golden_guards_the_dbr_cell_ordering(emu):
    """Negative control for golden dbr_f64_order: the same source built with MSIM_X_F32CELL=1 (the DBR calendar ordered
    by float32 time alone, as in round 1) must NOT reproduce it -- so the golden really exercises the exact float64
    order of python-typed instants inside one float32 spacing of the clock (examples/toc_dbr/simulation.py:200-289 on
    the SimPy heap)."""
    from manusim_b200.engine import decode_log

    g = load_golden("dbr_f64_order")
    eng = _engine(emu, g["scenario"], defines=("MSIM_X_F32CELL=1",))
    res = eng.run(1, tapes=[{k: g["tape_" + k] for k in "ABCD"}], n_log_reps=1, log_cap=len(g["ring"]) + 4096)
    ring, tm, val, _ = decode_log(res.log[0], int(res.log_count[0]))
    same = (int(res.n_events[0]) == g["steps"] and len(ring) == len(g["ring"])
            and np.array_equal(ring, g["ring"].astype(np.uint32)) and np.array_equal(tm, g["time"])
            and np.array_equal(val, g["value"]))
    assert not same


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name,status", list(zip(ERROR_NAMES, (1, 2))))
def test_emulated_kernel_value_error_cases(emu, name, status, kernel, monkeypatch):
    from manusim_b200.engine import decode_log

    monkeypatch.setenv("MSIM_NO_LEAN", "1" if kernel == "general" else "")
    g = load_golden(name)
    eng = _engine(emu, g["scenario"])
    res = eng.run(2, tapes=[{k: g["tape_" + k] for k in "ABCD"}] * 2, n_log_reps=2, log_cap=4096)
    assert (res.status == status).all()
    ring, tm, val, _ = decode_log(res.log[1], int(res.log_count[1]))
    assert np.array_equal(ring, g["ring"].astype(np.uint32))
    assert np.array_equal(tm, g["time"]) and np.array_equal(val, g["value"])


@pytest.mark.parametrize("seed", list(range(48)) if FULL else list(range(0, 48, 2)))
def test_emulated_kernel_random_scenarios_match_oracle(emu, seed):
    """Same scenarios and the same check as tests/test_gpu_fuzz.py (every 2nd seed unless MSIM_EMU_FULL=1)."""
    from test_gpu_fuzz import random_scenario

    from manusim_b200 import compile_scenario
    from manusim_b200.engine import decode_log
    from oracle.factory_oracle import run_oracle

    cfg, res, prod, policy, repair = random_scenario(seed)
    eng = emu.EmuEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair, max_production_orders=250,
                                         max_demand_orders=250, fifo_capacity=256))
    n = 3
    r = eng.run(n, seed=1000 + seed, n_log_reps=n, log_cap=1 << 16, record_tapes=1 << 14)
    cnt = r.rec["count"]
    assert cnt.max() < (1 << 14)
    for i in range(n):
        tp = {k: r.rec[k][i, : cnt[i, j]] for j, k in enumerate("ABCD")}
        o = run_oracle(cfg, res, prod, None, policy, tapes=tp, dbr_repair=repair)
        ring, tm, val, _ = decode_log(r.log[i], int(r.log_count[i]))
        if o["error"] is not None:
            assert r.status[i] == (1 if "Negative delay" in o["error"] else 2), (seed, i, r.status[i], o["error"])
        else:
            assert r.status[i] == 0, (seed, i, r.status[i])
            assert int(r.n_events[i]) == o["steps"], (seed, i)
        assert np.array_equal(ring, o["ring"].astype(np.uint32)), (seed, i, policy, cfg)
        assert np.array_equal(tm.view(np.uint32), o["time"].view(np.uint32)), (seed, i, policy, cfg)
        assert np.array_equal(val.view(np.uint32), o["value"].view(np.uint32)), (seed, i, policy, cfg)


@pytest.mark.parametrize("name", ["ss_lots", "dbr_limits", "js20_ondue"])
def test_host_buffer_entry_point_equals_device_buffer_entry_point(emu, name):
    """msim_run_host (scratch carving, staging of seeds / tapes / offsets, copy-back of accumulators and log rings in
    msim_api.cu) against msim_run on the same inputs: replay with MORE replications than logged ones, and Philox with
    explicit per-replication keys.  (The sanitizer run below repeats it under ASan.)"""
    g = load_golden(name)
    eng = _engine(emu, g["scenario"])
    tapes = [{k: g["tape_" + k] for k in "ABCD"}] * 3
    cap = len(g["ring"]) + 16
    a = eng.run(3, tapes=tapes, n_log_reps=2, log_cap=cap)
    b = eng.run_host(3, tapes=tapes, n_log_reps=2, log_cap=cap)
    for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts", "status", "log_count"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    n = int(a.log_count[0])
    assert np.array_equal(a.log[:, :n], b.log[:, :n])
    keys = np.array([11, 2**63 + 5, 12345678901234567], dtype=np.uint64)
    a = eng.run(3, rep_seeds=keys)
    b = eng.run_host(3, rep_seeds=keys)
    for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts", "status"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    c = eng.run_host(3, seed=99, rep_offset=1000)
    d = eng.run(3, seed=99, rep_offset=1000)
    assert np.array_equal(c.n_events, d.n_events) and np.array_equal(c.acc_sum, d.acc_sum)


@pytest.mark.parametrize("name", ["toy2_instantly", "ss_lots", "ss_queues", "js20_ondue", "dbr_limits"])
def test_lean_kernels_equal_the_general_ones(emu, name, monkeypatch):
    """A run that uses neither queue records nor tape recording launches the OPTS = 0 instantiations (without those
    branches); MSIM_NO_LEAN=1 forces the general kernel; ss_queues (log_queues) and the recording run below take the
    general kernel anyway."""
    from manusim_b200.engine import decode_log

    g = load_golden(name)
    eng = _engine(emu, g["scenario"])
    tapes = [{k: g["tape_" + k] for k in "ABCD"}] * 2
    cap = len(g["ring"]) + 16
    monkeypatch.setenv("MSIM_NO_LEAN", "1")
    ref = eng.run(2, tapes=tapes, n_log_reps=2, log_cap=cap)
    ref_p = eng.run(3, seed=21)
    monkeypatch.setenv("MSIM_NO_LEAN", "")
    r = eng.run(2, tapes=tapes, n_log_reps=2, log_cap=cap)
    assert (r.status == 0).all() and (r.n_events == g["steps"]).all()
    ring, tm, val, _ = decode_log(r.log[1], int(r.log_count[1]))
    assert np.array_equal(ring, g["ring"].astype(np.uint32)) and np.array_equal(tm, g["time"]) and \
        np.array_equal(val, g["value"])
    for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts"):
        assert np.array_equal(getattr(r, f), getattr(ref, f)), f
    lean_p = eng.run(3, seed=21)                              # Philox, lean
    rec_p = eng.run(3, seed=21, record_tapes=1 << 14)         # recording => general kernel
    for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts", "status"):
        assert np.array_equal(getattr(lean_p, f), getattr(ref_p, f)), f
        assert np.array_equal(getattr(rec_p, f), getattr(ref_p, f)), f
    assert rec_p.rec["count"].sum() > 0


def test_emulated_reduction_kernels_match_numpy(emu):
    """msim_reduce_metrics / msim_reduce_variables (block-level shuffles + __syncthreads) against plain numpy."""
    g = load_golden("ss_lots")
    eng = _engine(emu, g["scenario"])
    t = eng.tables
    n = 7
    res = eng.run(n, seed=5)
    assert (res.status == 0).all()
    res.status[3] = 5                      # a failed replication is skipped
    K, P, R = eng.K, t.P, t.R
    span = float(t.run_until) - float(t.warmup)
    ok = res.status == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = res.acc_sum / res.acc_cnt
    x = np.where(np.arange(K) < 3 * P, res.acc_sum,
                 np.where((np.arange(K) >= 11 * P) & (np.arange(K) < 11 * P + 3 * R), res.acc_sum / span, mean))
    x = x[ok]
    want = np.stack([np.nansum(x, 0), np.nansum(x * x, 0), np.sum(~np.isnan(x), 0)], 1)
    np.testing.assert_allclose(eng.reduce(res), want, rtol=1e-12, atol=1e-12)
    groups = [slice(v * P, (v + 1) * P) for v in range(11)] + \
             [slice(11 * P + v * R, 11 * P + (v + 1) * R) for v in range(4)] + \
             [slice(11 * P + 4 * R + v, 11 * P + 4 * R + v + 1) for v in range(3)]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        xv = np.stack([np.nanmean(x[:, s], axis=1) for s in groups], 1)
    wantv = np.stack([np.nansum(xv, 0), np.nansum(xv * xv, 0), np.sum(~np.isnan(xv), 0)], 1)
    np.testing.assert_allclose(eng.reduce(res, variables=True), wantv, rtol=1e-12, atol=1e-12)
    # msim_metric_values: the per-run numbers of metrics_*.csv; failed replication -> NaN row
    from manusim_b200.metrics import metric_values

    mv = eng.metric_values(res)
    want_mv = metric_values(res.acc_sum, res.acc_cnt, t, span)
    want_mv[~ok] = np.nan
    np.testing.assert_array_equal(mv, want_mv)


def test_kernel_source_is_clean_under_address_and_ub_sanitizers(emu):
    """All eight simulation-kernel instantiations and both reductions under ASan + UBSan, product CTA shape and one
    warp per CTA; then the negative control (an output buffer one element short must abort the child)."""
    if "thread" in _asan_build:
        _asan_build["thread"].join()
    if "error" in _asan_build:
        raise _asan_build["error"]
    emu.build("asan")
    env = emu.asan_env()
    script = str(HERE / "emu" / "sanitize_run.py")
    r = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-4000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["sanitize"] == "ok" and out["cases"] >= 26
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-4000:]
    bad = subprocess.run([sys.executable, script, "--inject"], env=env, capture_output=True, text=True, timeout=300)
    assert bad.returncode != 0 and "heap-buffer-overflow" in bad.stderr and "sim_kernel" in bad.stderr
    assert "INJECT NOT DETECTED" not in bad.stdout


def test_simt_shim_semantics_and_hang_detection(emu):
    """The shim itself: shuffle / redux / ballot / match.any results, block barrier, and the two failure modes it must
    report instead of hanging -- a collective that can never complete, lanes meeting at different primitives."""
    exe = emu.BUILD / "selftest"
    src = [str(emu.HERE / "selftest.cu"), str(emu.HERE / "emu_runtime.cpp")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I", str(emu.HERE / "include"),
           "-x", "c++"] + src + ["-o", str(exe), "-lpthread"]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert json.loads(r.stdout) == {"fails": 0, "divergence_detected": 1, "mismatch_detected": 1}
