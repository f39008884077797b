"""Diff oracle/minisimpy.py against the REAL SimPy package on every golden scenario (TEST INFRASTRUCTURE).

    ORACLE_SIMPY=real python -m oracle.check_simpy_pin          # needs `import simpy` to find the real package

The goldens under tests/golden were produced by the unmodified reference on minisimpy (oracle/gen_golden.py).  With the
real package selected this script re-runs (a) the oracle port on every golden scenario, seeded and in replay, and (b) --
when /root/reference is mounted -- the UNMODIFIED reference itself, and demands identical record streams, variate
tapes and `next(env._eid) - len(env._queue)` step counts.  Exit status 0 = minisimpy is pinned to the package on
everything the kernel's event order rests on; prints one JSON line per scenario.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main() -> int:
    from conftest import ERROR_NAMES, GOLDEN_NAMES, ORACLE_ONLY_NAMES, ORACLE_POLICY, load_golden

    from oracle import factory_oracle as fo
    from oracle import run_reference as rr

    if "minisimpy" in (getattr(fo.simpy, "__file__", "") or ""):
        print(json.dumps({"error": "the real simpy package was not selected (set ORACLE_SIMPY=real)"}))
        return 2
    bad = 0
    for name in GOLDEN_NAMES + ORACLE_ONLY_NAMES + ERROR_NAMES:
        g = load_golden(name)
        s = g["scenario"]
        row = {"scenario": name, "engine": fo.SIMPY_ENGINE}
        o = fo.run_oracle(s["cfg"], s["resources"], s["products"], s["seed"], ORACLE_POLICY[s["policy"]],
                          dbr_repair="P3" in s["patches"])
        keys = ("ring", "time", "value", "tape_A", "tape_B", "tape_C") + (() if name in ERROR_NAMES else ("tape_D",))
        row["port_identical"] = bool(all(np.array_equal(o[k].view(np.uint8), g[k].view(np.uint8)) for k in keys)
                                     and (name in ERROR_NAMES or o["steps"] == g["steps"])
                                     and (o["error"] or "") == g["error"])
        if rr.reference_available():
            r = rr.run_reference(s["cfg"], s["resources"], s["products"], s["seed"], policy=s["policy"],
                                 patches=s["patches"])
            row["reference_identical"] = bool(all(np.array_equal(r[k].view(np.uint8), g[k].view(np.uint8)) for k in keys)
                                              and (name in ERROR_NAMES or r["steps"] == g["steps"])
                                              and (r["error"] or "") == g["error"])
        bad += not (row["port_identical"] and row.get("reference_identical", True))
        print(json.dumps(row), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
