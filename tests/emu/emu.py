"""TEST INFRASTRUCTURE -- host-side SIMT emulation of the simulation kernel (see include/cuda_runtime.h here).

`build()` compiles the unmodified kernel sources with g++ against the shim headers into _build/libmsim_emu*.so;
`EmuEngine` drives that library through the same C ABI (include/msim.h) with numpy buffers, so the GPU parity tests
can be mirrored on CPU: sanitizer runs of the kernel source and bit-exact checks of event-order changes without
spending GPU minutes.  Nothing under manusim_b200/ imports this module or loads this library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
CSRC = ROOT / "manusim_b200" / "csrc"
BUILD = HERE / "_build"
SOURCES = [CSRC / "msim_kernel.cu", CSRC / "msim_reduce.cu", CSRC / "msim_api.cu", HERE / "emu_runtime.cpp"]
DEPS = SOURCES + [CSRC / "msim_device.cuh", CSRC / "philox.cuh", ROOT / "include" / "msim.h",
                  HERE / "include" / "cuda_runtime.h", HERE / "include" / "math_constants.h"]

FLAVOURS = {
    # -ffp-contract=off: the *_rn intrinsics are single IEEE operations (replay mode is bit-identical to the GPU)
    "plain": ["-O2", "-g"],
    # -O0 keeps the sanitizer build at ~20 s (eight kernel instantiations); the sanitizer runs use small scenarios
    # line-execution counts of the kernel source (gcov) for tools/instr_model.py
    "cov": ["-O0", "-g", "--coverage"],
    "asan": ["-O0", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"],
}


def _tag(flavour, defines):
    return flavour + "".join("_" + d.replace("=", "") for d in defines)


def lib_path(flavour: str = "plain", defines=()) -> Path:
    t = _tag(flavour, defines)
    return BUILD / ("libmsim_emu.so" if t == "plain" else f"libmsim_emu_{t}.so")


def build(flavour: str = "plain", force: bool = False, extra_defines=()) -> Path:
    """extra_defines: the kernel's experiment switches (e.g. ("MSIM_X_COLD=1",)) -> a separately named library"""
    extra_defines = tuple(extra_defines)
    out = lib_path(flavour, extra_defines)
    if not force and out.exists() and all(d.stat().st_mtime <= out.stat().st_mtime for d in DEPS):
        return out
    BUILD.mkdir(exist_ok=True)
    cmd = (["g++", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
            "-Wno-unused-function", "-Wno-maybe-uninitialized", "-I", str(HERE / "include")] + FLAVOURS[flavour]
           + [f"-D{d}" for d in extra_defines])
    objs, procs = [], []
    for s in SOURCES:                     # one g++ per translation unit, in parallel
        o = BUILD / f"{s.stem}_{_tag(flavour, extra_defines)}.o"
        objs.append(str(o))
        procs.append(subprocess.Popen(cmd + ["-c", "-x", "c++", str(s), "-o", str(o)], stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    errs = [p.communicate()[1] for p in procs]
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError("g++ failed on the emulated kernel build:\n" + "\n".join(errs)[-6000:])
    res = subprocess.run(cmd + objs + ["-o", str(out), "-lpthread"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link of the emulated kernel library failed:\n" + res.stderr[-6000:])
    return out


_libs = {}


def load(flavour: str = "plain", defines=()):
    key = _tag(flavour, tuple(defines))
    if key in _libs:
        return _libs[key]
    from manusim_b200 import _lib as L

    lib = C.CDLL(str(build(flavour, extra_defines=defines)))
    lib.msim_version.restype = C.c_int
    lib.msim_last_error.restype = C.c_char_p
    lib.msim_create.argtypes = [C.POINTER(L.Tables), C.POINTER(C.c_void_p)]
    lib.msim_destroy.argtypes = [C.c_void_p]
    lib.msim_query.argtypes = [C.c_void_p, C.POINTER(L.Info)]
    lib.msim_run.argtypes = [C.c_void_p, C.POINTER(L.RunArgs)]
    lib.msim_run_host.argtypes = [C.c_void_p, C.POINTER(L.HostArgs)]
    lib.msim_reduce_metrics.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p]
    lib.msim_reduce_variables.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p]
    lib.msim_metric_values.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p]
    _libs[key] = lib
    return lib


class EmuEngine:
    """numpy twin of manusim_b200.engine.BatchEngine.run for the emulated library ("device" memory = host memory)."""

    def __init__(self, tables, flavour: str = "plain", defines=()):
        from manusim_b200 import _lib as L
        from manusim_b200.engine import make_ctables

        self.L = L
        self.lib = load(flavour, defines)
        self.tables = tables
        ct, self._keep = make_ctables(tables)
        self.handle = C.c_void_p()
        self._check(self.lib.msim_create(C.byref(ct), C.byref(self.handle)))
        info = L.Info()
        self._check(self.lib.msim_query(self.handle, C.byref(info)))
        self.info = {f[0]: getattr(info, f[0]) for f in L.Info._fields_}
        self.K = info.n_rings

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.msim_last_error().decode()
            raise {-3: NotImplementedError, -1: ValueError}.get(rc, RuntimeError)(msg)

    def close(self):
        if self.handle and self.handle.value:
            self.lib.msim_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_reps, seed=0, rep_offset=0, tapes=None, rep_seeds=None, n_log_reps=0, log_cap=0, record_tapes=0):
        from manusim_b200.engine import pack_tapes

        K = self.K
        out = SimpleNamespace(
            acc_sum=np.full((n_reps, K), np.nan), acc_cnt=np.full((n_reps, K), 0x7FFFFFFF, dtype=np.uint32),
            n_events=np.zeros(n_reps, dtype=np.uint64), n_timeouts=np.zeros(n_reps, dtype=np.uint32),
            status=np.full(n_reps, 8, dtype=np.int32), log=None, log_count=None, rec=None)
        a = self.L.RunArgs()
        a.n_reps, a.rep_offset, a.philox_seed = n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF
        keep = []
        if tapes is not None:
            a.rng_mode = 1
            packed = tapes if isinstance(tapes, dict) else pack_tapes(tapes)
            for k in "ABCD":
                # exact-size copies: an out-of-bounds tape read is then visible to AddressSanitizer
                v = np.array(packed[k], copy=True)
                o = np.array(packed[k + "_off"], dtype=np.int64, copy=True)
                keep += [v, o]
                setattr(a, "tape_" + k, v.ctypes.data)
                setattr(a, f"tape_{k}_off", o.ctypes.data)
        else:
            a.rng_mode = 0
            if rep_seeds is not None:
                rs = np.ascontiguousarray(rep_seeds, dtype=np.uint64)
                keep.append(rs)
                a.rep_seeds = rs.ctypes.data
            if record_tapes > 0:
                rec = {"A": np.zeros((n_reps, record_tapes), dtype=np.float32),
                       "B": np.zeros((n_reps, record_tapes), dtype=np.float32),
                       "C": np.zeros((n_reps, record_tapes), dtype=np.float64),
                       "D": np.zeros((n_reps, record_tapes), dtype=np.float32),
                       "count": np.zeros((n_reps, 4), dtype=np.uint32)}
                a.rec_A, a.rec_B, a.rec_C, a.rec_D = (rec[k].ctypes.data for k in "ABCD")
                a.rec_cap, a.rec_count = record_tapes, rec["count"].ctypes.data
                out.rec = rec
        a.acc_sum, a.acc_cnt = out.acc_sum.ctypes.data, out.acc_cnt.ctypes.data
        a.n_events, a.n_timeouts, a.status = out.n_events.ctypes.data, out.n_timeouts.ctypes.data, out.status.ctypes.data
        if n_log_reps > 0 and log_cap > 0:
            out.log = np.zeros((min(n_log_reps, n_reps), log_cap, 4), dtype=np.int32)
            out.log_count = np.zeros(min(n_log_reps, n_reps), dtype=np.uint32)
            a.log, a.log_count = out.log.ctypes.data, out.log_count.ctypes.data
            a.log_cap, a.n_log_reps = log_cap, out.log.shape[0]
        self._check(self.lib.msim_run(self.handle, C.byref(a)))
        out._keep = keep
        return out

    def run_host(self, n_reps, seed=0, rep_offset=0, tapes=None, rep_seeds=None, n_log_reps=0, log_cap=0):
        """msim_run_host (host-buffer twin of msim_run: scratch carving, H2D staging, D2H copy-back in msim_api.cu)."""
        from manusim_b200.engine import pack_tapes

        K = self.K
        out = SimpleNamespace(
            acc_sum=np.full((n_reps, K), np.nan), acc_cnt=np.full((n_reps, K), 0x7FFFFFFF, dtype=np.uint32),
            n_events=np.zeros(n_reps, dtype=np.uint64), n_timeouts=np.zeros(n_reps, dtype=np.uint32),
            status=np.full(n_reps, 8, dtype=np.int32), log=None, log_count=None)
        a = self.L.HostArgs()
        a.n_reps, a.rep_offset, a.philox_seed = n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF
        keep = []
        if tapes is not None:
            a.rng_mode = 1
            packed = tapes if isinstance(tapes, dict) else pack_tapes(tapes)
            for k in "ABCD":
                v = np.array(packed[k], copy=True)
                o = np.array(packed[k + "_off"], dtype=np.int64, copy=True)
                if int(o[-1]) < len(v):
                    v = v[: max(int(o[-1]), 0)].copy() if int(o[-1]) > 0 else np.zeros(0, dtype=v.dtype)
                keep += [v, o]
                setattr(a, "tape_" + k, v.ctypes.data if len(v) else 0)
                setattr(a, f"tape_{k}_off", o.ctypes.data)
        else:
            a.rng_mode = 0
            if rep_seeds is not None:
                rs = np.ascontiguousarray(rep_seeds, dtype=np.uint64)
                keep.append(rs)
                a.rep_seeds = rs.ctypes.data
        a.acc_sum, a.acc_cnt = out.acc_sum.ctypes.data, out.acc_cnt.ctypes.data
        a.n_events, a.n_timeouts, a.status = out.n_events.ctypes.data, out.n_timeouts.ctypes.data, out.status.ctypes.data
        if n_log_reps > 0 and log_cap > 0:
            out.log = np.zeros((min(n_log_reps, n_reps), log_cap, 4), dtype=np.int32)
            out.log_count = np.zeros(min(n_log_reps, n_reps), dtype=np.uint32)
            a.log, a.log_count = out.log.ctypes.data, out.log_count.ctypes.data
            a.log_cap, a.n_log_reps = log_cap, out.log.shape[0]
        ms = C.c_double(-1.0)
        a.kernel_ms = C.pointer(ms)
        self._check(self.lib.msim_run_host(self.handle, C.byref(a)))
        out._keep = keep
        return out

    def reduce(self, res, variables=False):
        out = np.zeros((18 if variables else self.K, 3))
        fn = self.lib.msim_reduce_variables if variables else self.lib.msim_reduce_metrics
        self._check(fn(self.handle, res.acc_sum.ctypes.data, res.acc_cnt.ctypes.data, res.status.ctypes.data,
                       res.acc_sum.shape[0], out.ctypes.data, None))
        return out

    def metric_values(self, res):
        out = np.zeros((res.acc_sum.shape[0], self.K))
        self._check(self.lib.msim_metric_values(self.handle, res.acc_sum.ctypes.data, res.acc_cnt.ctypes.data,
                                                res.status.ctypes.data, res.acc_sum.shape[0], out.ctypes.data, None))
        return out


def asan_env():
    """Environment for a child python that dlopens the asan flavour."""
    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ)
    env["LD_PRELOAD"] = libasan
    env["ASAN_OPTIONS"] = "detect_leaks=0:abort_on_error=0:exitcode=66:detect_stack_use_after_return=0"
    env["UBSAN_OPTIONS"] = "print_stacktrace=1:halt_on_error=1"
    return env


if __name__ == "__main__":
    for fl in sys.argv[1:] or ["plain"]:
        print(build(fl, force=True))
