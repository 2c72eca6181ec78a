"""(mu, rho) initialisers -- host-side torch code, one-time, not on the hot path.

API parity with /root/reference/bayeformers/nn/parameters/initializations.py:14-60.
They deliberately consume torch's global generator in the reference's order
(mu first, then rho; quirk Q12) so a seeded conversion is reproducible against it.
"""
from __future__ import annotations

from typing import Tuple

from torch.nn import Parameter

Range = Tuple[float, float]


class Initialization:
    def __call__(self, mu: Parameter, rho: Parameter) -> Tuple[Parameter, Parameter]:
        raise NotImplementedError("Initialization not implemented yet")


class Uniform(Initialization):
    def __init__(self, mu_range: Range, rho_range: Range) -> None:
        self.mu_range, self.rho_range = mu_range, rho_range

    def __call__(self, mu: Parameter, rho: Parameter) -> Tuple[Parameter, Parameter]:
        lo, hi = self.mu_range
        mu.data = mu.data.uniform_(lo, hi)
        lo, hi = self.rho_range
        rho.data = rho.data.uniform_(lo, hi)
        return mu, rho


DEFAULT_UNIFORM = Uniform((-0.2, 0.2), (-5, -4))
