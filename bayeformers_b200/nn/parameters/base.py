"""Variational-parameter base types.

API parity with the reference's `bayeformers.nn.parameters.base`
(/root/reference/bayeformers/nn/parameters/base.py:17-68): the `parameter()`
factory, the `Parameter` interface (`sample`, `log_prob`) and the
`NoneParameter` stand-in used by bias-less layers.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Size, Tensor


def parameter(size: Size, dtype: torch.dtype = torch.float32) -> nn.Parameter:
    # fp32 masters regardless of `dtype`: the reference ignores the argument
    # (base.py:32, quirk Q4) and checkpoints depend on it.
    del dtype
    return nn.Parameter(torch.zeros(size, dtype=torch.float32))


class Parameter(nn.Module):
    """Interface of a distribution over a weight tensor."""

    def sample(self) -> Tensor:
        raise NotImplementedError("Sample not implemented yet")

    def log_prob(self, input: Tensor) -> Tensor:
        raise NotImplementedError("Log_prob not implemented yet")


class NoneParameter(Parameter):
    """Absent tensor (e.g. no bias): samples to None, contributes 0 to log-probs
    (reference base.py:55-68)."""

    def sample(self) -> None:
        return None

    def log_prob(self, input: Tensor) -> float:
        return 0.0
