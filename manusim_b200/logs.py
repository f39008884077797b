"""Host-side ``Logger``-compatible view over the record stream a replication streamed to its HBM ring.

Mirrors the observable surface of /root/reference/src/manusim/engine/logs.py (one attribute per variable holding
``dict[key] -> float32[mem_size, 2]``, ``log_index[variable][key]``, get_log/get_logs/get_variable_logs/
get_last_log_value/save_logs_to_file/save_all_logs/read_saved_logs/restart_log) so that code written against
``sim.logs`` keeps working.  The ring-wrap quirk (logs.py:59-66: the visible count restarts at 1 once a ring has
seen ``mem_size`` records, older rows are flushed to CSV only when a save path is set) is reproduced when the
stream is folded into the per-key rings.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import pandas as pd


class Logger:
    def __init__(self, logs_save_path=None, mem_size: int = 10000, enabled: bool = True, allocate_storage: bool = True):
        self.mem_size = mem_size
        self.enabled = enabled
        self.allocate_storage = allocate_storage
        self.log_index: Dict[str, Dict[str, int]] = {}
        self.logs_save_path = Path(logs_save_path) if logs_save_path else logs_save_path

    # ------------------------------------------------------------------ construction
    @classmethod
    def for_tables(cls, tables, logs_save_path=None, mem_size: int = 10000, enabled: bool = True):
        """Logger with the 18 standard variables of a compiled scenario (factory_sim.py:118-125)."""
        lg = cls(logs_save_path, mem_size, enabled)
        by_var = {}
        for variable, key in tables.ring_names():
            by_var.setdefault(variable, []).append(key)
        for variable, keys in by_var.items():
            lg.create_log(variable, keys)
        return lg

    def create_log(self, variable: str, keys):
        if not getattr(self, variable, False):
            var = {}
            self.log_index[variable] = {}
            for key in keys:
                var[key] = (np.zeros((self.mem_size, 2), dtype=np.float32) if self.allocate_storage
                            else np.empty((0, 2), dtype=np.float32))
                self.log_index[variable][key] = 0
            setattr(self, variable, var)

    def load_stream(self, ring_names, ring: np.ndarray, time: np.ndarray, value: np.ndarray):
        """Fold an ordered record stream into the rings exactly as successive ``log()`` calls would (logs.py:44-66)."""
        if not self.enabled:
            return
        order = np.argsort(ring, kind="stable")
        r_sorted = ring[order]
        bounds = np.searchsorted(r_sorted, np.arange(len(ring_names) + 1))
        m = self.mem_size
        for rid, (variable, key) in enumerate(ring_names):
            lo, hi = bounds[rid], bounds[rid + 1]
            n = hi - lo
            if n == 0:
                continue
            sel = order[lo:hi]
            recs = np.stack([time[sel], value[sel]], axis=1).astype(np.float32)
            li = ((n - 1) % m) + 1               # visible count after n log() calls
            c0 = n - li                          # records before the current cycle
            arr = getattr(self, variable)[key]
            if c0 > 0 and self.logs_save_path:
                # every completed cycle was appended to the CSV and the ring re-zeroed (save_logs_to_file + restart_log)
                self._append_csv(variable, key, recs[:c0])
                arr[:] = 0
            elif c0 > 0:
                # no save path: old rows stay in place behind the restarted index (silently dropped from metrics)
                prev = recs[c0 - m:c0]
                arr[li:] = prev[li:]
            arr[:li] = recs[c0:]
            self.log_index[variable][key] = int(li)

    # ------------------------------------------------------------------ reference API
    def log(self, variable: str, key: str, value: Tuple[float, float]):
        if not self.enabled:
            return
        var = getattr(self, variable, None)
        if var is None or variable not in self.log_index or key not in var:
            return
        index = self.log_index[variable][key] % self.mem_size
        if self.log_index[variable][key] >= self.mem_size and self.logs_save_path:
            self.save_logs_to_file(variable, key)
        var[key][index] = value
        self.log_index[variable][key] = index + 1

    def get_log(self, variable: str, key: str) -> np.ndarray:
        var = getattr(self, variable, None)
        if var is None:
            return np.empty((0, 2), dtype=np.float32)
        return var.get(key, np.empty((0, 2), dtype=np.float32))

    def get_last_log_value(self, variable: str, key: str):
        index = self.log_index.get(variable, {}).get(key, 0)
        if index == 0:
            return None
        var = getattr(self, variable, None)
        if var is None or key not in var:
            return None
        return tuple(var[key][(index - 1) % self.mem_size])

    def get_logs(self, variable: str, key: str) -> pd.DataFrame:
        var = getattr(self, variable, None)
        if var is None or key not in var:
            return pd.DataFrame(columns=["time", "value", "variable", "key"])
        df = pd.DataFrame(var[key][: self.log_index[variable][key]], columns=["time", "value"])
        df["variable"] = variable
        df["key"] = key
        return df

    def get_variable_logs(self, variable: str, saved_logs=False) -> pd.DataFrame:
        frames = []
        if saved_logs:
            saved = self.read_saved_logs(variable)
            if not saved.empty:
                frames.append(saved)
        for key in getattr(self, variable, {}).keys():
            frames.append(self.get_logs(variable, key))
        if frames:
            return pd.concat(frames, ignore_index=True)
        return pd.DataFrame(columns=["time", "value", "variable", "key"])

    def _append_csv(self, variable, key, rows):
        self.logs_save_path.mkdir(exist_ok=True, parents=True)
        path = Path(f"{self.logs_save_path}/log_{variable}_{key}.csv")
        df = pd.DataFrame(rows, columns=["time", "value"])
        df["variable"] = variable
        df["key"] = key
        df.to_csv(path, index=False, header=not path.exists(), mode="a")

    def save_logs_to_file(self, variable: str, key: str):
        if not self.logs_save_path:
            return
        if variable not in self.log_index or key not in self.log_index[variable]:
            return
        self._append_csv(variable, key, getattr(self, variable)[key][: self.log_index[variable][key]])
        self.restart_log(variable, key)

    def restart_log(self, variable: str, key: str):
        if variable not in self.log_index or key not in self.log_index[variable]:
            return
        self.log_index[variable][key] = 0
        getattr(self, variable)[key] = (np.zeros((self.mem_size, 2), dtype=np.float32) if self.allocate_storage
                                        else np.empty((0, 2), dtype=np.float32))

    def save_all_logs(self):
        for variable, keys in self.log_index.items():
            for key in keys.keys():
                self.save_logs_to_file(variable=variable, key=key)

    def read_saved_logs(self, variable: str) -> pd.DataFrame:
        if not self.logs_save_path:
            return pd.DataFrame()
        frames = []
        for path in list(Path(self.logs_save_path).rglob(f"**/*log_{variable}*.csv")):
            if path.is_file():
                df = pd.read_csv(path, low_memory=False)
                if not df.empty:
                    frames.append(df)
        return pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
