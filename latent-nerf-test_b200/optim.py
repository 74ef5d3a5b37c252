"""The step after the backward (SURVEY.md §8 f rank 4): the reference's optimiser on the texture,
``torch.optim.Adam(params, lr, betas=(0.9, 0.99), eps=1e-15)`` (reference ``src/latent_paint/training/trainer.py:93-95``,
``src/latent_paint_mesh/training/trainer.py:326-328``), as one CUDA kernel per parameter — optionally fed straight from
the texel-interleaved accumulation buffer of the vector-RED backward, which fuses the unpack into the update."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class FusedAdam:
    """Drop-in for ``torch.optim.Adam`` on fp32 CUDA parameters (no weight decay, no amsgrad, as the reference uses it):
    ``zero_grad`` / ``step`` with the same state names (``exp_avg``, ``exp_avg_sq``, ``step``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("lp_b200 FusedAdam: parameters must be contiguous fp32 CUDA tensors (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.state = {id(p): {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)} for p in self.params}

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _launch(self, p, st, grad=None, accum=None, channels=None, grad_out=None):
        st["step"] += 1
        a = _lib.LpAdamArgs()
        if accum is not None:
            a.accum = accum.data_ptr()
            a.grad = grad_out.data_ptr() if grad_out is not None else None
            a.C, a.ntex = int(channels), p.numel() // int(channels)
        else:
            a.grad = grad.data_ptr()
            a.C, a.ntex = 1, p.numel()
        a.param, a.exp_avg, a.exp_avg_sq = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
        a.lr, a.beta1, a.beta2, a.eps, a.step = self.lr, self.betas[0], self.betas[1], self.eps, st["step"]
        with torch.cuda.device(p.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
            _lib.check(_lib.lib().lp_adam_step(ctypes.byref(a), stream))

    @torch.no_grad()
    def step(self):
        for p in self.params:
            if p.grad is None:
                continue
            self._launch(p, self.state[id(p)], grad=p.grad.to(torch.float32).contiguous())

    @torch.no_grad()
    def step_from_accum(self, p, accum, grad_out=None):
        """Update the (1,C,T,T) texture ``p`` from the (T*T,4) interleaved accumulation buffer that
        ``lp_render_backward`` leaves with ``LP_FLAG_GRAD_INTERLEAVED`` (unpack fused into the update); ``grad_out``
        optionally receives the planar gradient."""
        C = p.shape[-3]
        if C > 4 or accum.numel() * accum.element_size() < 16 * (p.numel() // C):
            raise ValueError("step_from_accum: C <= 4 and an accumulation buffer of 16 bytes per texel are required")
        self._launch(p, self.state[id(p)], accum=accum, channels=C, grad_out=grad_out)
