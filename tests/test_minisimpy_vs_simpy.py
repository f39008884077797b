"""CPU: the automatic pin of the SimPy layer.  oracle/minisimpy.py restates simpy==4.1.1 (the reference's third-party
event kernel, /root/reference/requirements.txt:10) from memory; the package is not in this image's wheelhouse, so today
this test SKIPS.  The first time `import simpy` works -- an environment that has the wheel, or one unpacked under
baseline/_ref -- it re-runs every golden scenario on the real package (the oracle port, and the unmodified reference when
/root/reference is mounted) and demands the committed goldens bit for bit: record streams, variate tapes, step counts."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _real_simpy_importable():
    code = ("import sys; sys.path.insert(0, %r); from oracle.simpy_select import real_simpy; "
            "sys.exit(0 if real_simpy() is not None else 3)" % str(ROOT))
    return subprocess.run([sys.executable, "-c", code], capture_output=True).returncode == 0


def test_selector_defaults_to_the_restatement_and_never_aliases_it_as_the_package(monkeypatch):
    from oracle import simpy_select

    monkeypatch.delenv("ORACLE_SIMPY", raising=False)
    mod, what = simpy_select.pick()
    assert mod.__name__.endswith("minisimpy") and "unpinned" in what
    if not _real_simpy_importable():
        monkeypatch.setenv("ORACLE_SIMPY", "real")
        with pytest.raises(ImportError):
            simpy_select.pick()
        monkeypatch.setenv("ORACLE_SIMPY", "auto")
        assert simpy_select.pick()[0].__name__.endswith("minisimpy")


def test_minisimpy_reproduces_the_real_simpy_package_on_every_golden():
    if not _real_simpy_importable():
        pytest.skip("simpy==4.1.1 is not importable here (not in the offline wheelhouse): SimPy layer stays unpinned")
    r = subprocess.run([sys.executable, "-m", "oracle.check_simpy_pin"], cwd=ROOT, capture_output=True, text=True,
                       env={**os.environ, "ORACLE_SIMPY": "real"}, timeout=1800)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and rows, (r.stdout[-2000:], r.stderr[-2000:])
    assert all(row["port_identical"] and row.get("reference_identical", True) for row in rows), rows
