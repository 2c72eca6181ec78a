"""Short-sequence attention of the host model routed through the native kernels (opt-in).

The reference leaves attention to the host model (HuggingFace BERT calls torch's scaled_dot_product_attention between the
Bayesian query / key / value projections and the Bayesian output projection, examples/bert_glue.py).  With every Linear on
the tensor cores that library call became the largest non-contraction block of the training step at sequence length 128,
so `bayeformers_b200.accelerate_host_(model, attention=True)` registers an attention function with transformers'
`AttentionInterface` and points the model's config at it.  The function takes the case the kernels cover -- bf16, no
attention mask, not causal, T <= 128 (a multiple of 16), head width 64 -- and hands everything else to transformers' own
"sdpa" function, so results never depend on the switch beyond rounding and the dropout stream (Philox counter based,
regenerated in backward, instead of torch's generator).
"""
from __future__ import annotations

import torch

from ... import ops, runtime

NAME = "bayeformers_b200"
_registered = {"done": False}


def _fallback():
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    return ALL_ATTENTION_FUNCTIONS["sdpa"]


def attention_forward(module, query, key, value, attention_mask, dropout: float = 0.0, scaling=None, **kwargs):
    """transformers attention-interface signature: query / key / value [B, heads, T, d]; returns ([B, T, heads, d], None)."""
    usable = (attention_mask is None and not kwargs.get("is_causal", False) and not getattr(module, "is_causal", False)
              and ops.attention_supported(query, key, value))
    if not usable:
        return _fallback()(module, query, key, value, attention_mask, dropout=dropout, scaling=scaling, **kwargs)
    if not hasattr(module, "_bf_site"):
        module._bf_site, module._bf_calls = runtime.next_tensor_id(), 0
    module._bf_calls += 1
    drop = ops.DropoutSpec(p=float(dropout), seed=runtime.dropout_seed(), site_id=module._bf_site,
                           step=module._bf_calls & 0xFFFFFFFF)
    module._last_dropout = drop  # identity of this forward's mask (tests, debugging)
    scale = float(scaling) if scaling is not None else float(query.shape[-1]) ** -0.5
    # the boxes the pre-hook below handed to this forward's q / k / v projections: the backward kernel fills them with
    # the projections' bias gradients (column sums of dq, dk, dv per folded sample), consumed once
    boxes, module._bf_qkv_boxes = getattr(module, "_bf_qkv_boxes", None), None
    return ops.AttentionFn.apply(query, key, value, scale, drop, boxes, runtime.get_mc_samples()), None


def _qkv_pre_hook(module, args):
    """Before a self-attention block runs: give its Bayesian query / key / value projections a `bias_grad_box` each (see
    ops.BayesLinear: a filled box replaces the layer's own pass over its output gradient)."""
    boxes = []
    for name in ("query", "key", "value"):
        box = []
        getattr(module, name)._bias_grad_box = box
        boxes.append(box)
    module._bf_qkv_boxes = boxes


def use_native_attention_(model: torch.nn.Module, bias_grads: bool = False) -> int:
    """Point every transformers config found in `model` at the native attention function.  Returns how many configs
    were switched (0: not a transformers model, nothing done).  bias_grads: let the backward kernel emit the q / k / v
    projections' bias gradients (see accelerate_host_)."""
    try:
        from transformers import AttentionInterface
    except Exception:  # pragma: no cover
        return 0
    if not _registered["done"]:
        AttentionInterface.register(NAME, attention_forward)
        # transformers builds NO mask for attention names it does not know ("custom attention without equivalent mask
        # creation"): register torch-SDPA's mask builder under our name, so a padded batch still arrives with its
        # [B, 1, T, T] mask and is handed to the SDPA fallback instead of silently attending to the padding
        try:
            from transformers.masking_utils import AttentionMaskInterface, sdpa_mask
            AttentionMaskInterface.register(NAME, sdpa_mask)
        except Exception as e:  # pragma: no cover
            raise RuntimeError("this transformers version has no AttentionMaskInterface: the native attention cannot be "
                               "enabled safely (padded batches would lose their mask)") from e
        _registered["done"] = True
    seen, n = set(), 0
    for mod in model.modules():
        cfg = getattr(mod, "config", None)
        if cfg is None or id(cfg) in seen or not hasattr(cfg, "_attn_implementation"):
            continue
        seen.add(id(cfg))
        if getattr(cfg, "_attn_implementation", None) in ("sdpa", "eager", None, NAME):
            cfg._attn_implementation = NAME
            n += 1
    # dropout sites get their stream ids now (same construction order on every rank)
    from ..parameters.gaussian import Gaussian
    from .linear import Linear as BayesLinearModule
    for mod in model.modules():
        if type(mod).__name__.endswith("SelfAttention") and not hasattr(mod, "_bf_site"):
            mod._bf_site, mod._bf_calls = runtime.next_tensor_id(), 0
            qkv = [getattr(mod, name, None) for name in ("query", "key", "value")]
            if bias_grads and all(isinstance(l, BayesLinearModule) and isinstance(l.bias, Gaussian) for l in qkv):
                mod.register_forward_pre_hook(_qkv_pre_hook)
    return n
