"""Fused global-norm clipping + AdamW.

    opt = bf.optim.ClipAdamW(params, lr=2e-5, weight_decay=0.01, max_grad_norm=1.0)
    loss.backward(); opt.step(); opt.zero_grad()

does in two launches what the reference's loop does with
`clip_grad_norm_(params, 1.0); AdamW.step()` (examples/bert_glue.py:240-241):
one pass for the global gradient norm, one pass that clips and updates (the
separate "scale the gradients" pass disappears).  Arithmetic is
torch.optim.AdamW's; bf16 parameters (what `cast_frequentist_` makes of the host
model's embeddings / LayerNorms) get fp32 moments AND an fp32 master copy that the
update runs on -- at lr 2e-5 a step is smaller than half a bf16 ulp of a 0.02-sized
weight and would otherwise round away.  The
per-tensor step counts live on the device (advanced by the kernel), so `step()`
can be captured in a CUDA graph.
CUDA-only, like the rest of the hot path.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional

import torch

from . import _lib, ops
from ._lib import BF_BF16, BF_F32, BfOptDesc


class ClipAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None,
                 model: Optional[torch.nn.Module] = None) -> None:
        """model: when given, every `Gaussian` of it whose rho is among `params` gets a sigma cache that this optimizer
        keeps current (softplus of the updated rho is written next to the update), which the multi-tensor sampling
        kernel then reads instead of recomputing softplus per element (nn/parameters/gaussian.py)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("ClipAdamW got no trainable parameters")
        for p in self.params:
            ops._require_cuda(p, "parameter")
            if p.dtype not in (torch.float32, torch.bfloat16):
                raise TypeError(f"unsupported parameter dtype {p.dtype}")
            if not p.is_contiguous():
                raise ValueError("parameters must be contiguous")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = -1.0 if max_grad_norm is None else float(max_grad_norm)
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("ClipAdamW: all parameters must live on one device (one process per GPU)")
        self.device = dev
        # fp32 masters of the bf16 parameters (None for fp32 parameters, which are their own masters)
        self.master = [p.detach().float().clone() if p.dtype == torch.bfloat16 else None for p in self.params]
        # sigma caches of the rho tensors (None for everything else)
        self.gaussians = [None] * len(self.params)
        if model is not None:
            from .nn.parameters.gaussian import Gaussian
            by_rho = {id(g.rho): g for g in model.modules() if isinstance(g, Gaussian)}
            for i, p in enumerate(self.params):
                g = by_rho.get(id(p))
                if g is not None and p.dtype == torch.float32:
                    g.refresh_sigma_cache()
                    self.gaussians[i] = g
        self.exp_avg = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in self.params]
        self.exp_avg_sq = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in self.params]
        self.step_count = torch.zeros(len(self.params), dtype=torch.float32, device=dev)  # per tensor, like torch
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)  # of the last step, before clipping
        lib = _lib.load()
        ce = lib.bf_optim_chunk_elems()
        chunks: List[int] = []
        for ti, p in enumerate(self.params):
            for c in range(max((p.numel() + ce - 1) // ce, 1)):
                chunks += [ti, c]
        self.n_chunks = len(chunks) // 2
        self.d_chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.d_ws = torch.empty(lib.bf_clip_adamw_workspace_bytes(self.n_chunks), dtype=torch.uint8, device=dev)
        # descriptor table: a small ring of pinned staging buffers (gradient pointers change every step in eager mode)
        self._nbytes = ctypes.sizeof(BfOptDesc) * len(self.params)
        self._ring = [torch.empty(self._nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        self._events = [None] * len(self._ring)
        self._slot = 0
        self.d_descs = torch.empty(self._nbytes, dtype=torch.uint8, device=dev)
        self._last_sig = None

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _upload_descs(self) -> None:
        sig = tuple(0 if p.grad is None else p.grad.data_ptr() for p in self.params) + tuple(p.data_ptr() for p in self.params)
        if sig == self._last_sig:
            return
        descs = (BfOptDesc * len(self.params))()
        for i, p in enumerate(self.params):
            g = p.grad
            d = descs[i]
            d.param, d.exp_avg, d.exp_avg_sq = p.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr()
            d.n, d.dtype = p.numel(), (BF_BF16 if p.dtype == torch.bfloat16 else BF_F32)
            d.master = None if self.master[i] is None else self.master[i].data_ptr()
            gs = self.gaussians[i]
            d.sigma_out = None if (gs is None or getattr(gs, "_sigma", None) is None) else gs._sigma.data_ptr()
            if g is None:
                d.grad = None
            else:
                if g.dtype != p.dtype or not g.is_contiguous():
                    g = g.to(p.dtype).contiguous()
                    p.grad = g
                d.grad = g.data_ptr()
            al = 8 if p.dtype == torch.bfloat16 else 16
            d.vec = int(p.numel() % 4 == 0 and p.data_ptr() % al == 0 and (g is None or g.data_ptr() % al == 0))
        slot = self._slot
        self._slot = (slot + 1) % len(self._ring)
        if self._events[slot] is not None:
            self._events[slot].synchronize()  # the copy that last used this staging buffer has completed
        host = self._ring[slot]
        ctypes.memmove(host.data_ptr(), ctypes.addressof(descs), self._nbytes)
        self.d_descs.copy_(host, non_blocking=True)
        if not torch.cuda.is_current_stream_capturing():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._events[slot] = ev
        self._last_sig = sig

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """One update; returns the (device) global gradient norm before clipping."""
        lib = _lib.load()
        with ops.on_device(self.device):
            self._upload_descs()
            nbytes = float(sum(p.numel() * (2 * p.element_size() + 16 + p.element_size()) for p in self.params))
            rc = ops._timed("clip_adamw", nbytes, self.device, lambda: lib.bf_clip_adamw_step(
                self.d_descs.data_ptr(), self.d_chunks.data_ptr(), self.n_chunks, self.lr, self.betas[0], self.betas[1],
                self.eps, self.weight_decay, self.max_grad_norm, self.step_count.data_ptr(), self.grad_norm.data_ptr(),
                self.d_ws.data_ptr(), ops._stream(self.device)))
        _lib.check(rc, "bf_clip_adamw_step")
        ops.stats["launches"] += 2
        for p, gs in zip(self.params, self.gaussians):
            # the kernel has just written softplus(updated rho) for every tensor that had a gradient: if torch-side
            # code had invalidated the cache in between (load_state_dict, ...), it is current again from here on
            if gs is not None and p.grad is not None and gs.sigma_cache() is None and getattr(gs, "_sigma", None) is not None:
                gs.refresh_sigma_cache(written_by_kernel=True)
        return self.grad_norm

    # ---- checkpointing (the reference saves no optimizer state, bert_glue.py:303-309; a resumable run needs it)
    def state_dict(self) -> dict:
        return {"exp_avg": [t.detach().cpu() for t in self.exp_avg],
                "exp_avg_sq": [t.detach().cpu() for t in self.exp_avg_sq],
                "master": [None if t is None else t.detach().cpu() for t in self.master],
                "step_count": self.step_count.detach().cpu(),
                "hyper": {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                          "max_grad_norm": self.max_grad_norm}}

    @torch.no_grad()
    def load_state_dict(self, state: dict) -> None:
        if len(state["exp_avg"]) != len(self.params):
            raise ValueError("optimizer state does not match the parameter list")
        for dst, src in zip(self.exp_avg, state["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, state["exp_avg_sq"]):
            dst.copy_(src)
        for p, dst, src in zip(self.params, self.master, state["master"]):
            if dst is not None:
                dst.copy_(p.detach().float() if src is None else src)
        self.step_count.copy_(state["step_count"])
        h = state.get("hyper", {})
        self.lr, self.eps = float(h.get("lr", self.lr)), float(h.get("eps", self.eps))
        self.betas = tuple(h.get("betas", self.betas))
        self.weight_decay = float(h.get("weight_decay", self.weight_decay))
        self.max_grad_norm = float(h.get("max_grad_norm", self.max_grad_norm))
