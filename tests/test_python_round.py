"""CPU: the kernel's `round(python float, 6)` (DBR utilisation accounting on python-typed instants, reference
factory_sim.py:443) against CPython's own round() on adversarial inputs -- products that land on or next to a decimal
half-way point, where rint(x*1e6)/1e6 alone disagrees with CPython on about one input in eight.  The function body is
taken from the .cu sources as they are and compiled for the host."""
import ctypes
import math
import random
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_device_pyround6_equals_cpython_round(tmp_path, source="msim_kernel.cu"):
    text = (ROOT / "manusim_b200" / "csrc" / source).read_text()
    m = re.search(r"__device__ __forceinline__ double pyround6\(double x\) \{.*?\n\}", text, re.S)
    assert m, "pyround6 not found"
    body = m.group(0).replace("__device__ __forceinline__", 'extern "C"')
    src = "#include <math.h>\n#define __dmul_rn(a, b) ((a) * (b))\n#define __ddiv_rn(a, b) ((a) / (b))\n" + body + "\n"
    (tmp_path / "r.cpp").write_text(src)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(tmp_path / "r.cpp"), "-o",
                    str(tmp_path / "r.so")], check=True)
    f = ctypes.CDLL(str(tmp_path / "r.so")).pyround6
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.c_double]
    rng = random.Random(7)
    naive_differs = 0
    for _ in range(60000):
        k = rng.randrange(0, 10 ** 9)
        x = (k + 0.5) / 1e6
        for xx in (x, math.nextafter(x, 0.0), math.nextafter(x, 1e9), rng.uniform(0.0, 500.0), float(k) / 1e6):
            want = round(xx, 6)
            assert f(xx) == want, (xx, f(xx), want)
            naive_differs += (math.floor(xx * 1e6 + 0.5) / 1e6 != want) or (round(xx * 1e6) / 1e6 != want)
    assert naive_differs > 1000          # the inputs are adversarial: the one-line formula fails on many of them
