"""`bnn.Model`: wrapper that exposes the model-level ELBO scalars.

API parity with /root/reference/bayeformers/nn/model.py:16-89: duck-typed
detection of Bayesian children (both attribute names present), pass-through
`forward` that calls `self.model.forward` directly (no hooks, quirk Q8), and
`log_prior()` / `log_variational_posterior()` METHODS returning the sum of the
children's scalars of the most recent forward, accumulated in module order.
"""
from __future__ import annotations

import warnings
from typing import Any, List, Optional

import torch
from torch import Tensor
from torch.nn import Module


def is_module_bayesian(module: Module) -> bool:
    return hasattr(module, "log_prior") and hasattr(module, "log_variational_posterior")


class Model(Module):
    def __init__(self, model: Optional[Module] = None) -> None:
        super().__init__()
        self.model = model
        self._presampler = None  # set by bayeformers_b200.enable_presample

    def forward(self, *args, **kwargs) -> Any:
        if self.model is None:
            raise NotImplementedError("Forward pass not implemented yet")
        from .. import runtime
        S = runtime.get_mc_samples()
        runtime.reset_sinks()  # gradient sinks are per forward
        if self._presampler is not None:
            # extension: draw every Bayesian Linear's weights for this forward in one multi-tensor launch
            self._presampler.run(S)
        if S > 1:
            # folded forward: remember S*B so Bayesian layers can expand broadcast inputs (leading dimension 1)
            lead = next((int(v.shape[0]) for v in list(args) + list(kwargs.values())
                         if torch.is_tensor(v) and v.dim() >= 1), None)
            runtime.set_folded_rows(lead)
            try:
                return self.model.forward(*args, **kwargs)
            finally:
                runtime.set_folded_rows(None)
        return self.model.forward(*args, **kwargs)

    @property
    def bayesian_children(self) -> List[Module]:
        return [m for m in self.modules() if is_module_bayesian(m) and m is not self]

    def _total(self, name: str) -> Tensor:
        children = self.bayesian_children
        if not children:
            warnings.warn("No Bayesian Child is present in this model")
        value = 0.0
        for child in children:
            # grad-carrying value when the layer ran with kl_grad=True, else the
            # reference's detached Parameter
            live = getattr(child, "live_" + name, None)
            if live is None:  # per-sample [S] values of a folded forward, else the 0-dim Parameter
                live = getattr(child, name + "_samples", None)
            value = value + (live if live is not None else getattr(child, name))
        return value

    def log_prior(self) -> Tensor:
        return self._total("log_prior")

    def log_variational_posterior(self) -> Tensor:
        return self._total("log_variational_posterior")
