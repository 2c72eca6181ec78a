"""S-sample helpers for user training / evaluation loops.

The reference's examples carry their own `sample_bayesian` loops
(/root/reference/examples/bert_glue.py:56-73, examples/bert_squad.py:190-212):
S sequential forwards, stacking the logits and the two model scalars, then
means over the sample axis.  `sample_bayesian` here returns the same tuple
layout but runs the S samples as ONE folded forward (`mc_samples(S)`), which is
bit-identical given the same eps and lets every
Bayesian layer do a single batched contraction.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple, Union

import torch

from . import runtime
from .nn.model import Model

Select = Union[int, str, Sequence[Union[int, str]], Callable]


def _fold(value, S: int):
    """Repeat a batched input along dim 0 so that rows [s*B, (s+1)*B) are sample s's copy."""
    if torch.is_tensor(value) and value.dim() >= 1:
        return value.repeat(S, *([1] * (value.dim() - 1)))
    return value


def _pick(outputs, select: Select):
    if callable(select):
        return select(outputs)
    if isinstance(select, (list, tuple)):
        return tuple(_pick(outputs, s) for s in select)
    if isinstance(select, str):
        return outputs[select] if isinstance(outputs, dict) else getattr(outputs, select)
    return outputs[select]


def sample_bayesian(model: Model, inputs: Dict[str, torch.Tensor], samples: int, select: Select = 0,
                    fold: bool = True) -> Tuple:
    """Run `samples` Monte-Carlo forwards of `model(**inputs)`.

    select   which model output(s) are the logits: an index / attribute name / key, a
             sequence of those (e.g. (-2, -1) for start/end logits), or a callable.
    fold     True: one folded forward of S*B rows; False: the reference's sequential
             S-loop (same results given the same eps; slower).

    Returns `(raw_logits, logits, log_prior, log_variational_posterior)` as the
    reference's helper does: raw_logits `[S, B, ...]` (per-sample predictions, what
    predictive-uncertainty metrics such as `acc_std` are computed from,
    bert_glue.py:186), logits = their mean over S, and the two scalars averaged
    over samples.  With a sequence `select`, raw_logits / logits are tuples.
    """
    S = int(samples)
    if S < 1:
        raise ValueError("samples must be >= 1")
    if fold:
        folded = {k: _fold(v, S) for k, v in inputs.items()}
        with runtime.mc_samples(S):
            out = _pick(model(**folded), select)
        lp, lq = model.log_prior(), model.log_variational_posterior()

        def unfold(t):
            return t.reshape(S, t.shape[0] // S, *t.shape[1:])

        raw = tuple(unfold(t) for t in out) if isinstance(out, tuple) else unfold(out)
        log_prior, log_q = lp.mean(), lq.mean()
    else:
        outs, lps, lqs = [], [], []
        for _ in range(S):
            outs.append(_pick(model(**inputs), select))
            lps.append(model.log_prior())
            lqs.append(model.log_variational_posterior())
        raw = (tuple(torch.stack([o[i] for o in outs]) for i in range(len(outs[0])))
               if isinstance(outs[0], tuple) else torch.stack(outs))
        log_prior, log_q = torch.stack(lps).mean(), torch.stack(lqs).mean()
    mean = tuple(r.mean(0) for r in raw) if isinstance(raw, tuple) else raw.mean(0)
    return raw, mean, log_prior, log_q


def predictive_stats(raw_logits: torch.Tensor, labels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Per-example predictive uncertainty from per-sample logits `[S, B, C]`: mean
    class probabilities, predictive entropy, and (with labels) the mean / std over
    samples of the accuracy -- the `acc` / `acc_std` the reference logs
    (examples/bert_glue.py:176-190)."""
    probs = raw_logits.float().softmax(-1)
    mean_p = probs.mean(0)
    out = {"probs": mean_p, "entropy": -(mean_p * mean_p.clamp_min(1e-30).log()).sum(-1),
           "prediction": mean_p.argmax(-1)}
    if labels is not None:
        acc_s = (raw_logits.argmax(-1) == labels.unsqueeze(0)).float().mean(1)
        out["acc"], out["acc_std"] = acc_s.mean(), acc_s.std(unbiased=False)
    return out
