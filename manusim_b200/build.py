"""Build libmsim.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmsim.so"
SOURCES = ["msim_kernel.cu", "msim_reduce.cu", "msim_api.cu"]
HEADERS = ["msim_device.cuh", "philox.cuh", "../../include/msim.h"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS]
    return any(d.stat().st_mtime > t for d in deps)


def variant_path(tag: str) -> Path:
    """A/B builds (tools/ab_probe.py): same sources, extra -D switches, loaded with MSIM_LIB=<path>."""
    return PKG / f"libmsim_{tag}.so"


def build_lib(force: bool = False, verbose: bool = False, defines=(), tag: str = "") -> Path:
    """One nvcc per translation unit (in parallel), then one link: libmsim.so holds the simulation kernel
    (msim_kernel.cu), the experiment-statistics reductions (msim_reduce.cu) and the C ABI (msim_api.cu).
    `tag` + `defines` build an experiment variant next to it (libmsim_<tag>.so)."""
    LIB = variant_path(tag) if tag else globals()["LIB"]
    if not force and not tag and not needs_build():
        return LIB
    nvcc = _nvcc()
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
             "-Xptxas", "-v" if verbose else "-O3"] + [f"-D{d}" for d in defines]
    objdir = PKG / "build" / (tag or "default")
    objdir.mkdir(parents=True, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        o = objdir / (Path(s).stem + ".o")
        objs.append(str(o))
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", str(CSRC / s), "-o", str(o)], stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate() for p in procs]
    if verbose:
        for _, err in outs:
            sys.stderr.write(err)
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError("nvcc failed:\n" + "\n".join(o + e for o, e in outs))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    build_lib(force=True, verbose=True)
    print(LIB)
