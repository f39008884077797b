"""Stub of omegaconf.DictConfig (reference metrics.py:10, examples/*/simulation.py): a plain dict subclass."""


class DictConfig(dict):
    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as exc:
            raise AttributeError(item) from exc


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True):
        return cfg
