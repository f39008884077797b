"""Stub of hydra.types.RunMode (reference experiment.py:10,19)."""
import enum


class RunMode(enum.Enum):
    RUN = 1
    MULTIRUN = 2
