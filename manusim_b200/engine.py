"""BatchEngine: Python host of the batched replication kernel (thin layer over the C ABI in include/msim.h).

PyTorch is used only for device/pinned memory, streams and torch.distributed; all simulation work happens in
libmsim.so.  No CPU fallback: constructing an engine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .compiler import ScenarioTables

RNG_PHILOX, RNG_REPLAY = 0, 1

STATUS_TEXT = {0: "ok", 1: "Negative delay (SimPy ValueError)", 2: "Container amount <= 0 (SimPy ValueError)",
               3: "production-order pool overflow", 4: "demand-order pool overflow", 5: "event FIFO overflow",
               6: "replay tape exhausted", 7: "WIP get blocked (unsupported)", 8: "not run"}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def rep_keys(seed: int, global_index) -> np.ndarray:
    """Philox replication keys, same mapping as msim::rep_key (csrc/philox.cuh): a function of the experiment
    seed and the GLOBAL replication index, hence independent of how replications are sharded."""
    with np.errstate(over="ignore"):
        idx = np.asarray(global_index).astype(np.uint64)
        return _splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ _splitmix64(idx + np.uint64(0x632BE59BD9B4E019)))


def _dist_array(dists):
    arr = (_lib.Dist * len(dists))()
    for i, d in enumerate(dists):
        arr[i].kind = d.kind
        arr[i].raw0, arr[i].raw1, arr[i].d0, arr[i].d1 = d.raw0, d.raw1, d.d0, d.d1
    return arr


def pack_tapes(tapes: Sequence[Dict[str, np.ndarray]]):
    """[{A,B,C,D}] per replication -> concatenated value arrays + int64 offsets (C-ABI replay layout)."""
    out = {}
    for k, dt in (("A", np.float32), ("B", np.float32), ("C", np.float64), ("D", np.float32)):
        vals = [np.asarray(t[k], dtype=dt) for t in tapes]
        off = np.zeros(len(tapes) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(v) for v in vals])
        out[k] = np.concatenate(vals) if vals and off[-1] > 0 else np.zeros(1, dtype=dt)
        out[k + "_off"] = off
    return out


def make_ctables(t: ScenarioTables):
    """ScenarioTables -> (msim_tables_t, list of the numpy/ctypes arrays its pointers refer to)."""
    keep = []
    ct = _lib.Tables()
    ct.n_resources, ct.n_products, ct.n_steps = t.R, t.P, int(t.route_start[-1])
    rs = np.ascontiguousarray(t.route_start, dtype=np.int32)
    sr = np.ascontiguousarray(t.step_resource, dtype=np.int32)
    sb = np.ascontiguousarray(t.shipping_buffer, dtype=np.float64)
    keep += [rs, sr, sb]
    ct.route_start = rs.ctypes.data_as(C.POINTER(C.c_int32))
    ct.step_resource = sr.ctypes.data_as(C.POINTER(C.c_int32))
    ct.prod_shipping_buffer = sb.ctypes.data_as(C.POINTER(C.c_double))
    for name in ("step_time", "res_setup", "res_tbf", "res_ttr", "prod_freq", "prod_qty", "prod_due"):
        arr = _dist_array(getattr(t, name))
        keep.append(arr)
        setattr(ct, name, arr)
    ct.run_until, ct.warmup = float(t.run_until), float(t.warmup)
    ct.delivery_mode, ct.policy, ct.log_queues = t.delivery_mode, t.policy, int(t.log_queues)
    if t.dbr:
        d = t.dbr
        cpt = np.ascontiguousarray(d["ccr_pt"], dtype=np.float64)
        keep.append(cpt)
        ct.dbr_ccr, ct.dbr_repair = d["ccr"], int(d["repair"])
        ct.dbr_interval, ct.dbr_cb_target = float(d["interval"]), d["cb_target"]
        ct.dbr_release_limit, ct.dbr_ccr_limit = d["release_limit"], d["ccr_limit"]
        ct.dbr_min_lot, ct.dbr_ccr_setup = d["min_lot"], d["ccr_setup"]
        ct.dbr_ccr_pt = cpt.ctypes.data_as(C.POINTER(C.c_double))
    ct.max_production_orders, ct.max_demand_orders = t.max_production_orders, t.max_demand_orders
    ct.fifo_capacity = t.fifo_capacity
    return ct, keep


@dataclass
class BatchResult:
    acc_sum: "object"       # [n, K] float64
    acc_cnt: "object"       # [n, K] uint32 (torch: int32 view)
    n_events: "object"      # [n] SimPy-equivalent events
    n_timeouts: "object"    # [n] logical events (calendar pops)
    status: "object"        # [n]
    log: Optional["object"] = None        # [n_log, log_cap, 4] int32 view of 128-bit records
    log_count: Optional["object"] = None  # [n_log]
    rec: Optional[dict] = None            # recorded tapes (Philox mode)
    kernel_ms: Optional[float] = None


class BatchEngine:
    """Compiled scenario on one GPU."""

    def __init__(self, tables: ScenarioTables, device: Optional[int] = None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("manusim_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.tables = tables
        ct, self._keep = make_ctables(tables)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msim_create(C.byref(ct), C.byref(self.handle)))
        info = _lib.Info()
        _lib.check(self.lib.msim_query(self.handle, C.byref(info)))
        self.info = {f[0]: getattr(info, f[0]) for f in _lib.Info._fields_}
        self.K = info.n_rings

    @property
    def resident_replications(self) -> int:
        """Replications the GPU simulates at once (one per resident warp).  Kernel time is quantised in waves of this many:
        a batch that is a multiple of it wastes nothing, a half-filled last wave costs 0.7 of a whole one."""
        return int(self.info["warps_per_block"] * self.info["blocks_per_sm"] * self.info["n_sms"])

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.msim_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device-buffer path
    def run(self, n_reps: int, seed: int = 0, rep_offset: int = 0, tapes: Optional[Sequence[dict]] = None,
            rep_seeds=None, n_log_reps: int = 0, log_cap: int = 0, record_tapes: int = 0, stream=None,
            out: Optional[BatchResult] = None) -> BatchResult:
        """Launch n_reps replications; inputs/outputs are device tensors; asynchronous on `stream`."""
        torch = self.torch
        dev = self.device
        K = self.K
        if out is None:
            out = self.alloc(n_reps, n_log_reps, log_cap)
        a = _lib.RunArgs()
        a.n_reps, a.rep_offset, a.philox_seed = n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF
        keep = []
        if tapes is not None:
            a.rng_mode = RNG_REPLAY
            packed = tapes if isinstance(tapes, dict) else pack_tapes(tapes)
            for k in ("A", "B", "C", "D"):
                v = packed[k] if torch.is_tensor(packed[k]) else torch.from_numpy(packed[k]).to(dev)
                o = packed[k + "_off"] if torch.is_tensor(packed[k + "_off"]) else torch.from_numpy(packed[k + "_off"]).to(dev)
                keep += [v, o]
                setattr(a, "tape_" + k, v.data_ptr())
                setattr(a, f"tape_{k}_off", o.data_ptr())
        else:
            a.rng_mode = RNG_PHILOX
            if rep_seeds is not None:
                rs = rep_seeds if torch.is_tensor(rep_seeds) else torch.from_numpy(
                    np.asarray(rep_seeds, dtype=np.uint64).view(np.int64)).to(dev)
                keep.append(rs)
                a.rep_seeds = rs.data_ptr()
            if record_tapes > 0:
                rec = {"A": torch.zeros((n_reps, record_tapes), dtype=torch.float32, device=dev),
                       "B": torch.zeros((n_reps, record_tapes), dtype=torch.float32, device=dev),
                       "C": torch.zeros((n_reps, record_tapes), dtype=torch.float64, device=dev),
                       "D": torch.zeros((n_reps, record_tapes), dtype=torch.float32, device=dev),
                       "count": torch.zeros((n_reps, 4), dtype=torch.int32, device=dev)}
                a.rec_A, a.rec_B, a.rec_C, a.rec_D = (rec[k].data_ptr() for k in "ABCD")
                a.rec_cap, a.rec_count = record_tapes, rec["count"].data_ptr()
                out.rec = rec
        a.acc_sum, a.acc_cnt = out.acc_sum.data_ptr(), out.acc_cnt.data_ptr()
        a.n_events, a.n_timeouts, a.status = out.n_events.data_ptr(), out.n_timeouts.data_ptr(), out.status.data_ptr()
        if out.log is not None:
            a.log, a.log_count = out.log.data_ptr(), out.log_count.data_ptr()
            a.log_cap, a.n_log_reps = out.log.shape[1], out.log.shape[0]
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        a.stream = s.cuda_stream
        with torch.cuda.device(dev):
            _lib.check(self.lib.msim_run(self.handle, C.byref(a)))
        out._keep = keep
        return out

    def retry_overflow(self, res: BatchResult, seed: int = 0, rep_offset: int = 0, rep_seeds=None) -> int:
        """Replications that ran out of shared-memory pool slots (status 3/4/5: the factory got more congested than
        the configured pools) are re-simulated with the largest pools on a twin engine and patched into `res`.
        Philox mode only (keys are per replication, so the re-run reproduces the same trajectory).  Synchronises.
        Returns the number of replications re-run."""
        return self.retry_collect(self.retry_launch(res, seed, rep_offset, rep_seeds))

    def retry_launch(self, res: BatchResult, seed: int = 0, rep_offset: int = 0, rep_seeds=None, stream=None):
        """First half of retry_overflow: waits for `res`, launches the re-run (on `stream`) and returns a context for
        retry_collect, or None when nothing overflowed.  Lets callers overlap the re-runs of several engines."""
        torch = self.torch
        torch.cuda.synchronize(self.device)
        st = res.status
        bad = torch.nonzero((st >= 3) & (st <= 5)).flatten()
        if bad.numel() == 0:
            return None
        if getattr(self, "_big", None) is None:
            import copy

            t = copy.copy(self.tables)
            t.max_production_orders, t.max_demand_orders, t.fifo_capacity = 250, 250, 256
            self._big = BatchEngine(t, self.device.index)
        idx = bad.cpu().numpy()
        if rep_seeds is not None:
            rs = rep_seeds.cpu().numpy() if torch.is_tensor(rep_seeds) else np.asarray(rep_seeds)
            keys = rs.view(np.uint64)[idx] if rs.dtype != np.uint64 else rs[idx]
        else:
            keys = rep_keys(seed, idx.astype(np.int64) + rep_offset)
        n_log = int((idx < (res.log.shape[0] if res.log is not None else 0)).sum())
        import contextlib

        with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
            # the key upload and the output allocation must be ordered on the same stream as the kernel
            sub = self._big.run(len(idx), rep_seeds=np.ascontiguousarray(keys),
                                n_log_reps=len(idx) if n_log else 0, log_cap=res.log.shape[1] if n_log else 0,
                                stream=stream)
        return (res, bad, idx, sub, n_log)

    def retry_collect(self, ctx) -> int:
        if ctx is None:
            return 0
        torch = self.torch
        res, bad, idx, sub, n_log = ctx
        torch.cuda.synchronize(self.device)
        for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts", "status"):
            getattr(res, f)[bad] = getattr(sub, f)
        if n_log:
            for j, i in enumerate(idx):
                if i < res.log.shape[0]:
                    res.log[i] = sub.log[j]
                    res.log_count[i] = sub.log_count[j]
        return int(bad.numel())

    def alloc(self, n_reps: int, n_log_reps: int = 0, log_cap: int = 0) -> BatchResult:
        torch = self.torch
        dev = self.device
        res = BatchResult(
            acc_sum=torch.empty((n_reps, self.K), dtype=torch.float64, device=dev),
            acc_cnt=torch.empty((n_reps, self.K), dtype=torch.int32, device=dev),
            n_events=torch.empty((n_reps,), dtype=torch.int64, device=dev),
            n_timeouts=torch.empty((n_reps,), dtype=torch.int32, device=dev),
            status=torch.full((n_reps,), 8, dtype=torch.int32, device=dev))
        if n_log_reps > 0 and log_cap > 0:
            res.log = torch.empty((min(n_log_reps, n_reps), log_cap, 4), dtype=torch.int32, device=dev)
            res.log_count = torch.zeros((min(n_log_reps, n_reps),), dtype=torch.int32, device=dev)
        return res

    def reduce(self, res: BatchResult, out=None, stream=None):
        """Per-ring {sum x, sum x^2, n} over this shard's replications -> [K,3] float64 (accumulated into out)."""
        torch = self.torch
        if out is None:
            out = torch.zeros((self.K, 3), dtype=torch.float64, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msim_reduce_metrics(self.handle, res.acc_sum.data_ptr(), res.acc_cnt.data_ptr(),
                                                    res.status.data_ptr(), res.acc_sum.shape[0], out.data_ptr(),
                                                    s.cuda_stream))
        return out

    def reduce_variables(self, res: BatchResult, out=None, stream=None):
        """Per-variable {sum x, sum x^2, n} with x = the replication's mean over the variable's keys -> [18,3]."""
        torch = self.torch
        if out is None:
            out = torch.zeros((18, 3), dtype=torch.float64, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msim_reduce_variables(self.handle, res.acc_sum.data_ptr(), res.acc_cnt.data_ptr(),
                                                      res.status.data_ptr(), res.acc_sum.shape[0], out.data_ptr(),
                                                      s.cuda_stream))
        return out

    def metric_values(self, res: BatchResult, out=None, stream=None):
        """Per-replication metric values [n, K] float64 on the device (NaN: empty mean-metric or failed replication)."""
        torch = self.torch
        n = res.acc_sum.shape[0]
        if out is None:
            out = torch.empty((n, self.K), dtype=torch.float64, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msim_metric_values(self.handle, res.acc_sum.data_ptr(), res.acc_cnt.data_ptr(),
                                                   res.status.data_ptr(), n, out.data_ptr(), s.cuda_stream))
        return out

    # ------------------------------------------------------------------ host-buffer path (end to end)
    def alloc_host(self, n_reps: int, n_log_reps: int = 0, log_cap: int = 0) -> BatchResult:
        torch = self.torch
        pin = dict(pin_memory=True)
        res = BatchResult(
            acc_sum=torch.empty((n_reps, self.K), dtype=torch.float64, **pin),
            acc_cnt=torch.empty((n_reps, self.K), dtype=torch.int32, **pin),
            n_events=torch.empty((n_reps,), dtype=torch.int64, **pin),
            n_timeouts=torch.empty((n_reps,), dtype=torch.int32, **pin),
            status=torch.empty((n_reps,), dtype=torch.int32, **pin))
        if n_log_reps > 0 and log_cap > 0:
            res.log = torch.empty((min(n_log_reps, n_reps), log_cap, 4), dtype=torch.int32, **pin)
            res.log_count = torch.zeros((min(n_log_reps, n_reps),), dtype=torch.int32, **pin)
        return res

    def run_host(self, n_reps: int, seed: int = 0, rep_offset: int = 0, tapes=None, rep_seeds=None,
                 out: Optional[BatchResult] = None, n_log_reps: int = 0, log_cap: int = 0,
                 retry_overflow: bool = True, device_log: Optional[BatchResult] = None, stats=None) -> BatchResult:
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call, synchronous).  Replications that
        overflowed the shared-memory pools are re-run on the big-pool twin engine (Philox mode) unless disabled.
        `device_log`: a device BatchResult (from alloc) whose record rings the replications stream into -- they stay
        in HBM; `stats`: device [K,3] float64 tensor the call accumulates the per-ring {sum x, sum x^2, n} into (the
        vector the ranks allreduce; the reduction skips failed replications, so the re-run of overflowed ones simply
        adds its own contribution)."""
        torch = self.torch
        if out is None:
            out = self.alloc_host(n_reps, n_log_reps, log_cap)
        a = _lib.HostArgs()
        a.n_reps, a.rep_offset, a.philox_seed = n_reps, rep_offset, seed & 0xFFFFFFFFFFFFFFFF
        keep = []
        if tapes is not None:
            a.rng_mode = RNG_REPLAY
            packed = tapes if isinstance(tapes, dict) else pack_tapes(tapes)
            for k in ("A", "B", "C", "D"):
                v, o = packed[k], packed[k + "_off"]
                keep += [v, o]
                setattr(a, "tape_" + k, v.data_ptr() if torch.is_tensor(v) else v.ctypes.data)
                setattr(a, f"tape_{k}_off", o.data_ptr() if torch.is_tensor(o) else o.ctypes.data)
        else:
            a.rng_mode = RNG_PHILOX
            if rep_seeds is not None:
                rs = rep_seeds if torch.is_tensor(rep_seeds) else np.ascontiguousarray(rep_seeds, dtype=np.uint64)
                keep.append(rs)
                a.rep_seeds = rs.data_ptr() if torch.is_tensor(rs) else rs.ctypes.data
        a.acc_sum, a.acc_cnt = out.acc_sum.data_ptr(), out.acc_cnt.data_ptr()
        a.n_events, a.n_timeouts, a.status = out.n_events.data_ptr(), out.n_timeouts.data_ptr(), out.status.data_ptr()
        if out.log is not None:
            a.log, a.log_count = out.log.data_ptr(), out.log_count.data_ptr()
            a.log_cap, a.n_log_reps = out.log.shape[1], out.log.shape[0]
        ms = C.c_double(0.0)
        a.kernel_ms = C.pointer(ms)
        if device_log is not None and device_log.log is not None:
            if out.log is not None:
                raise ValueError("host log rings and device_log are mutually exclusive")
            a.log_dev, a.log_count_dev = device_log.log.data_ptr(), device_log.log_count.data_ptr()
            a.log_cap_dev, a.n_log_reps_dev = device_log.log.shape[1], device_log.log.shape[0]
        if stats is not None:
            a.stats_dev = stats.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.msim_run_host(self.handle, C.byref(a)))
        out.kernel_ms = ms.value
        if retry_overflow and tapes is None:
            st = out.status.numpy()
            bad = np.nonzero((st >= 3) & (st <= 5))[0]
            if len(bad):
                if getattr(self, "_big", None) is None:
                    import copy

                    t = copy.copy(self.tables)
                    t.max_production_orders, t.max_demand_orders, t.fifo_capacity = 250, 250, 256
                    self._big = BatchEngine(t, self.device.index)
                if rep_seeds is not None:
                    rs = rep_seeds.numpy() if torch.is_tensor(rep_seeds) else np.asarray(rep_seeds)
                    keys = rs.view(np.uint64)[bad]
                else:
                    keys = rep_keys(seed, bad.astype(np.int64) + rep_offset)
                sub = self._big.run_host(len(bad), rep_seeds=np.ascontiguousarray(keys), retry_overflow=False, stats=stats)
                idx = torch.from_numpy(bad)
                for f in ("acc_sum", "acc_cnt", "n_events", "n_timeouts", "status"):
                    getattr(out, f)[idx] = getattr(sub, f)
                out.kernel_ms += sub.kernel_ms
        return out


def decode_log(log_row: np.ndarray, count: int):
    """[cap,4] int32 view of 128-bit records -> (ring u32, time f32, value f32, seq u32) in emission order."""
    cap = log_row.shape[0]
    raw = np.ascontiguousarray(log_row).view(np.uint32).reshape(cap, 4)
    if count > cap:   # ring wrapped: oldest record sits right after the newest
        start = count % cap
        raw = np.concatenate([raw[start:], raw[:start]])
        n = cap
    else:
        n = count
    raw = raw[:n]
    return (raw[:, 2].copy(), raw[:, 0].copy().view(np.float32), raw[:, 1].copy().view(np.float32), raw[:, 3].copy())
