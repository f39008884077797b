"""Import stub (TEST INFRASTRUCTURE): lets /root/reference example scripts import without hydra-core 1.3.2.
Only the names the reference touches at import time exist (examples/*/simulation.py: ``@hydra.main``)."""


def main(*_args, **_kwargs):
    def deco(fn):
        return fn
    return deco
