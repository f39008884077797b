"""TEST INFRASTRUCTURE: the emulated library behind the BatchEngine surface (torch CPU tensors in / out), so that the
host logic above the C ABI -- FactorySimulation.run_simulation, ExperimentRunner (run tree, experiment-level files,
failure handling, log-ring re-runs), sharding + allreduce under gloo -- runs in the CPU suite on the same kernel
source the GPU executes.  Never imported by the product."""
from __future__ import annotations

import numpy as np
import torch

import emu
from manusim_b200.engine import BatchResult


def _t(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a)


class EmuTorchEngine:
    def __init__(self, tables, defines=()):
        self._e = emu.EmuEngine(tables, defines=defines)
        self.tables = tables
        self.K = self._e.K
        self.info = self._e.info
        self.device = torch.device("cpu")
        self.launches = []                    # (n_reps, n_log_reps, log_cap) per run() call, for the tests

    def synchronize(self):
        pass

    def close(self):
        self._e.close()

    def run(self, n_reps, seed=0, rep_offset=0, tapes=None, rep_seeds=None, n_log_reps=0, log_cap=0, record_tapes=0,
            stream=None, out=None):
        if torch.is_tensor(rep_seeds):
            rep_seeds = rep_seeds.numpy().view(np.uint64)
        r = self._e.run(n_reps, seed=seed, rep_offset=rep_offset, tapes=tapes, rep_seeds=rep_seeds,
                        n_log_reps=n_log_reps, log_cap=log_cap, record_tapes=record_tapes)
        self.launches.append((n_reps, n_log_reps, log_cap))
        res = BatchResult(acc_sum=_t(r.acc_sum), acc_cnt=_t(r.acc_cnt), n_events=_t(r.n_events),
                          n_timeouts=_t(r.n_timeouts), status=_t(r.status), log=_t(r.log), log_count=_t(r.log_count))
        res._raw = r
        return res

    def retry_overflow(self, res, seed=0, rep_offset=0, rep_seeds=None):
        return 0

    def _ns(self, res):
        from types import SimpleNamespace

        return SimpleNamespace(acc_sum=res.acc_sum.numpy(), acc_cnt=res.acc_cnt.numpy().view(np.uint32),
                               status=res.status.numpy())

    def reduce(self, res, out=None, stream=None):
        m = torch.from_numpy(self._e.reduce(self._ns(res)))
        if out is not None:
            out += m
            return out
        return m

    def reduce_variables(self, res, out=None, stream=None):
        m = torch.from_numpy(self._e.reduce(self._ns(res), variables=True))
        if out is not None:
            out += m
            return out
        return m

    def metric_values(self, res, out=None, stream=None):
        return torch.from_numpy(self._e.metric_values(self._ns(res)))
