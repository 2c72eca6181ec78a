"""`bayeformers_b200.nn` -- same export list as the reference's `bayeformers.nn`
(/root/reference/bayeformers/nn/__init__.py:3-25) plus the two layers the
north star adds (Embedding, LayerNorm)."""
import torch.nn as nn

from .layers.embedding import Embedding
from .layers.layernorm import LayerNorm
from .layers.linear import Linear
from .model import Model
from .parameters.base import NoneParameter, Parameter
from .parameters.gaussian import DEFAULT_SCALED_GAUSSIAN_MIXTURE, Gaussian, ScaledGaussianMixture
from .parameters.initializations import DEFAULT_UNIFORM, Initialization, Uniform

# frequentist class -> Bayesian replacement used by to_bayesian.  The reference
# registers nn.Linear only; Embedding / LayerNorm are opt-in through
# to_bayesian(..., layers=...) so the default conversion stays identical.
TORCH2BAYE = {nn.Linear: Linear}
TORCH2BAYE_ALL = {nn.Linear: Linear, nn.Embedding: Embedding, nn.LayerNorm: LayerNorm}

__all__ = ["Linear", "Embedding", "LayerNorm", "Model", "NoneParameter", "Parameter",
           "DEFAULT_SCALED_GAUSSIAN_MIXTURE", "Gaussian", "ScaledGaussianMixture", "DEFAULT_UNIFORM",
           "Initialization", "Uniform", "TORCH2BAYE", "TORCH2BAYE_ALL"]
