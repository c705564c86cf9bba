"""Import-only stub (oracle infrastructure): satisfies `from omegaconf import ...` in src/config.py."""


class DictConfig:
    pass


class OmegaConf:
    pass


def open_dict(x):
    pass
