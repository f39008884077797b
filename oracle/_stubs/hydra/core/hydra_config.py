"""Stub of hydra.core.hydra_config.HydraConfig (reference experiment.py:9,19,121-123)."""


class HydraConfig:
    @staticmethod
    def initialized():
        return False

    @staticmethod
    def get():
        raise RuntimeError("HydraConfig stub: no hydra run is active; pass save_folder_path explicitly")
