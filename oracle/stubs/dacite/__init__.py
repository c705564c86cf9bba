"""Import-only stub (oracle infrastructure): satisfies `from dacite import from_dict` in src/config.py."""


def from_dict(*a, **k):
    raise NotImplementedError("dacite stub: not used on the forward path")
