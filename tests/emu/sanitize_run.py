"""TEST INFRASTRUCTURE -- drive the AddressSanitizer/UBSan flavour of the emulated kernel library (run as a child
process with LD_PRELOAD=libasan, see emu.asan_env()).  Every kernel instantiation (4 policies x replay/Philox) and both
reduction kernels run on small scenarios, once with the product's CTA shape and once with one warp per CTA so that each
replication slab is its own allocation.  `--inject` shortens one output buffer by one element (negative control: the
sanitizer must stop the process)."""
import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))

import emu  # noqa: E402
from conftest import load_golden  # noqa: E402

from manusim_b200 import compile_scenario, scenarios  # noqa: E402
from manusim_b200.engine import decode_log  # noqa: E402

POLICY = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}


def replay(name, n=2):
    g = load_golden(name)
    scn = g["scenario"]
    t = compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy=POLICY[scn["policy"]],
                         dbr_repair="P3" in scn["patches"])
    eng = emu.EmuEngine(t, "asan")
    res = eng.run(n, tapes=[{k: g["tape_" + k] for k in "ABCD"}] * n, n_log_reps=n, log_cap=len(g["ring"]) + 8)
    if g["error"] in ("", "None"):
        assert (res.status == 0).all() and (res.n_events == g["steps"]).all(), (name, res.status)
    ring, tm, val, _ = decode_log(res.log[n - 1], int(res.log_count[n - 1]))
    assert np.array_equal(ring, g["ring"].astype(np.uint32)) and np.array_equal(tm, g["time"]) and \
        np.array_equal(val, g["value"]), name
    eng.reduce(res)
    eng.reduce(res, variables=True)
    # host-buffer entry point: scratch carving + staging + copy-back (msim_api.cu) under the sanitizers
    h = eng.run_host(n + 1, tapes=[{k: g["tape_" + k] for k in "ABCD"}] * (n + 1), n_log_reps=n, log_cap=len(g["ring"]) + 8)
    assert np.array_equal(h.n_events[:n], res.n_events) and np.array_equal(h.log_count, res.log_count), name
    h = eng.run_host(3, seed=5)
    assert h.status.shape == (3,)
    eng.close()


def philox(policy, mode, repair=False, log_queues=False):
    if policy == "dbr":
        sim, res, prod, _ = scenarios.load_example("toc_dbr")
        cfg = dict(sim, run_until=300, warmup=50)
        cfg.pop("print_mode", None)
    else:
        cfg, res, prod = scenarios.jobshop20(run_until=150, warmup=30, mode=mode)
        for i, r in enumerate(res.values()):
            r["tbf"] = {"dist": "expo", "params": [40 + 5 * i]}
    cfg["log_queues"] = log_queues
    eng = emu.EmuEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=repair), "asan")
    n = 5                                   # more replications than one CTA holds with MSIM_EMU_WPB=1
    r = eng.run(n, seed=99, n_log_reps=n - 1, log_cap=256, record_tapes=4096)   # small ring: wraps
    assert (r.status == 0).all(), (policy, mode, r.status)
    r2 = eng.run(n, seed=99)                # no logs, no recording: the fast paths
    assert np.array_equal(r2.n_events, r.n_events) and np.array_equal(r2.acc_cnt, r.acc_cnt)
    eng.close()


def main():
    inject = "--inject" in sys.argv
    cases = 0
    # (CTA shape, general kernel forced): the product's CTA shape with the kernels a run picks by itself (lean when it
    # uses neither queue records nor tape recording), then one warp per CTA -- every slab is its own allocation -- with the
    # general (OPTS = 3) kernels forced for every run
    configs = [("", ""), ("1", "1")]
    for wpb, general in () if inject else configs:
        os.environ["MSIM_NO_LEAN"] = general
        if wpb:
            os.environ["MSIM_EMU_WPB"] = wpb
        else:
            os.environ.pop("MSIM_EMU_WPB", None)
        for name in ("toy2", "toy2_instantly", "ss_lots", "ss_queues", "dbr_limits", "err_negdelay", "err_amount"):
            replay(name)
            cases += 1
        for policy, mode in (("fifo", "asReady"), ("fifo", "onDue"), ("max_priority", "instantly"), ("edd", "onDue")):
            philox(policy, mode, log_queues=(policy == "edd"))
            cases += 1
        philox("dbr", "instantly", repair=True)
        philox("dbr", "instantly", repair=False)
        cases += 2
    os.environ.pop("MSIM_NO_LEAN", None)
    if inject:
        g = load_golden("toy2")
        scn = g["scenario"]
        eng = emu.EmuEngine(compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy="fifo"), "asan")
        a = eng.L.RunArgs()
        import ctypes as C
        n, K = 2, eng.K
        bufs = dict(acc_sum=np.zeros(n * K), acc_cnt=np.zeros(n * K - 1, dtype=np.uint32),     # one element short
                    n_events=np.zeros(n, dtype=np.uint64), status=np.zeros(n, dtype=np.int32))
        a.n_reps, a.rng_mode = n, 0
        for k, v in bufs.items():
            setattr(a, k, v.ctypes.data)
        eng.lib.msim_run(eng.handle, C.byref(a))
        print("INJECT NOT DETECTED")
        return
    print(json.dumps({"sanitize": "ok", "cases": cases}))


if __name__ == "__main__":
    main()
