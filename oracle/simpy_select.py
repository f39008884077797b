"""Which SimPy the oracle runs on (TEST INFRASTRUCTURE).

Default: ``oracle/minisimpy.py``, the from-memory restatement of SimPy 4.1.1 (the third-party package the reference
pins in /root/reference/requirements.txt:10 -- not vendored, not installable offline).  ``ORACLE_SIMPY=real`` selects
the real package instead, importable either from the environment or from ``baseline/_ref`` (where a driver-provided
wheel would be unpacked); ``ORACLE_SIMPY=auto`` takes the real package when it is importable and the restatement
otherwise.  tests/test_minisimpy_vs_simpy.py uses this to diff the two the day a wheel exists; bench.py's CPU arm uses
``auto`` and says which engine it timed.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def real_simpy():
    """The real simpy package or None."""
    ref = ROOT / "baseline" / "_ref"
    if ref.is_dir() and str(ref) not in sys.path:
        sys.path.append(str(ref))
    mod = sys.modules.get("simpy")
    if mod is not None and getattr(mod, "__name__", "") == "simpy" and hasattr(mod, "Environment") \
            and "minisimpy" not in (getattr(mod, "__file__", "") or ""):
        return mod
    try:
        saved = sys.modules.pop("simpy", None)       # a minisimpy alias registered by run_reference must not shadow it
        mod = importlib.import_module("simpy")
        if "minisimpy" in (getattr(mod, "__file__", "") or ""):
            raise ImportError
        return mod
    except ImportError:
        if saved is not None:
            sys.modules["simpy"] = saved
        return None


def pick():
    """(module, description) according to ORACLE_SIMPY = mini (default) | real | auto."""
    mode = os.environ.get("ORACLE_SIMPY", "mini")
    if mode in ("real", "auto"):
        mod = real_simpy()
        if mod is not None:
            return mod, f"real SimPy {getattr(mod, '__version__', '?')}"
        if mode == "real":
            raise ImportError("ORACLE_SIMPY=real but the simpy package is not importable")
    from oracle import minisimpy

    return minisimpy, "oracle/minisimpy.py (SimPy 4.1.1 restated from memory; parity unpinned at this layer)"
