"""Gaussians container with the reference's attribute-bag semantics (src/utils/gaussians_types.py:4-38)."""
from __future__ import annotations

from torch import Tensor


class Gaussians:
    def __init__(self, means=None, covariances=None, harmonics=None, opacities=None, scales=None, rotations=None, **kwargs):
        self.means: Tensor = means
        self.covariances: Tensor = covariances
        self.harmonics: Tensor = harmonics
        self.opacities: Tensor = opacities
        self.scales: Tensor = scales
        self.rotations: Tensor = rotations
        for k, v in kwargs.items():
            setattr(self, k, v)

    def detach_cpu_copy(self) -> "Gaussians":
        out = Gaussians()
        for name, value in vars(self).items():
            setattr(out, name, value.detach().cpu() if isinstance(value, Tensor) else value)
        return out
