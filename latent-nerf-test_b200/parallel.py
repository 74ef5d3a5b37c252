"""View-sharded data parallelism for the render path.

The reference is single-process; its unit of work is a camera view.  Views are independent, so
rank r renders the contiguous block ``shard_views(n, r, world)`` of a batch against a replicated
mesh / texture, scatter-adds into its own texture gradient, and ONE all-reduce(sum) of that
gradient per step makes every rank's optimiser step identical to the single-GPU step over the
whole batch (up to fp32 summation order).  One process per GPU, ``torch.distributed`` (NCCL on
GPUs over NVLink; gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world_size: int) -> tuple[int, int]:
    """Half-open view range of ``rank``: blocks differ by at most one view, earlier ranks get the
    larger blocks; empty when there are more ranks than views."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(num_views, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradientBucket:
    """Flat fp32 buffer the learnable render inputs' gradients live in (texture first, then e.g.
    the background-sphere colours), so a step needs a single in-place all-reduce and no staging
    copy: ``param.grad`` tensors are views into the bucket."""

    def __init__(self, params):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("GradientBucket needs at least one parameter")
        device = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None, async_op=False):
        """Sum the bucket over all ranks (no-op without an initialised process group)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def render_views_sharded(render_fn, view_params: dict, num_views: int, group=None):
    """Call ``render_fn(**shard)`` on this rank's slice of every per-view tensor in
    ``view_params`` (tensors whose first dimension is ``num_views``; everything else is passed
    through).  Returns ``(outputs, (lo, hi))``; ``outputs`` is None for an empty shard."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_views(num_views, rank, world)
    if hi == lo:
        return None, (lo, hi)
    shard = {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == num_views else v)
             for k, v in view_params.items()}
    return render_fn(**shard), (lo, hi)


class SymmetricGradientBuffer:
    """Flat fp32 gradient buffer in symmetric memory + the library's own all-reduce kernels over NVLink /
    NVSwitch peer memory (``lp_allreduce_multimem`` — in-switch reduction — or the two-shot
    ``lp_allreduce_p2p``).  ``torch.distributed._symmetric_memory`` is the plumbing: allocation, rendezvous
    (peer / multicast mappings) and the stream-ordered barriers around the kernels.

    ``flat[:numel]`` is the payload; the tail pads the length to a multiple of ``4 * world`` floats so every
    rank owns an equally sized, 16-byte aligned slice.
    """

    def __init__(self, numel: int, device, group=None, interleaved_texels: int = 0, channels: int = 0,
                 with_params: bool = False):
        import ctypes

        import torch.distributed._symmetric_memory as symm

        from . import _lib
        self._lib, self._ctypes = _lib, ctypes
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        quantum = 4 * self.world
        self.numel = numel
        self.padded = (numel + quantum - 1) // quantum * quantum
        # fused form (``interleaved_texels`` = Th * Tw of ONE texture with ``channels`` <= 4 planes, numel =
        # channels * texels): the same allocation also holds the (Th,Tw,4) accumulation buffer the vector-RED
        # backward scatters into (LP_FLAG_GRAD_INTERLEAVED); lp_allreduce_unpack then reduces, unpacks and
        # broadcasts in one kernel
        self.texels, self.channels = int(interleaved_texels), int(channels)
        self.fused = bool(self.texels) and self.texels % quantum == 0 and 0 < self.channels <= 4 \
            and numel == self.channels * self.texels
        self.accum_floats = 4 * self.texels if self.fused else 0
        # layout of the allocation (floats): planar gradient | accumulation buffer | [planar parameters] | flag block
        self.param_floats = self.padded if (with_params and self.fused) else 0
        flag_floats = _lib.LP_EXCHANGE_FLAG_BYTES // 4
        self.flat_all = symm.empty(self.padded + self.accum_floats + self.param_floats + flag_floats, dtype=torch.float32, device=device)
        self.flat_all.zero_()
        self.flat = self.flat_all[:self.padded]
        self.accum = self.flat_all[self.padded:self.padded + self.accum_floats] if self.fused else None      # (texels, 4) interleaved
        self.accum_offset = 4 * self.padded                                   # bytes from the allocation's base
        self.param_offset = 4 * (self.padded + self.accum_floats)
        self.params = self.flat_all[self.padded + self.accum_floats:self.padded + self.accum_floats + self.param_floats] \
            if self.param_floats else None                                    # planar (C, texels): the texture itself
        self.flags_offset = 4 * (self.padded + self.accum_floats + self.param_floats)
        self.adam_state = None
        self.handle = symm.rendezvous(self.flat_all, self.group.group_name)
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.mode = "multimem" if self.multicast_ptr else "p2p"
        self.use_fused = self.fused          # callers may switch back to unpack + all-reduce on the planar gradient
        self.one_launch = True               # fused form: handshakes inside the kernel (False: barrier, kernel, barrier)

    def view(self, shape):
        n = 1
        for s in shape:
            n *= s
        return self.flat[:n].view(shape)

    def exchange_step(self, adam=None):
        """The exchange as ONE kernel launch (``lp_exchange_step``): the handshakes between the ranks happen inside the
        kernel through the flag block of this allocation, no host-enqueued barrier.  Needs the fused layout.
        ``adam = dict(lr, betas, eps)`` switches on the sharded optimiser epilogue: every rank updates its slice of
        ``self.params`` (allocate with ``with_params=True``) and broadcasts it; the optimiser state is sharded."""
        if not self.fused:
            raise RuntimeError("exchange_step needs the interleaved accumulation buffer in the allocation")
        L, c = self._lib.lib(), self._ctypes
        a = self._lib.LpExchangeArgs()
        a.multicast_base = c.c_void_p(self.multicast_ptr) if self.mode == "multimem" else None
        a.buffer_ptrs_dev = c.c_void_p(self.handle.buffer_ptrs_dev)
        a.accum_offset, a.grad_offset, a.flags_offset = self.accum_offset, 0, self.flags_offset
        a.ntex, a.C, a.rank, a.world = self.texels, self.channels, self.rank, self.world
        if adam is not None:
            if self.params is None:
                raise RuntimeError("the optimiser epilogue needs with_params=True")
            if self.adam_state is None:
                n = self.channels * self.texels // self.world
                self.adam_state = {"step": 0, "exp_avg": torch.zeros(n, device=self.flat.device),
                                   "exp_avg_sq": torch.zeros(n, device=self.flat.device)}
            st = self.adam_state
            st["step"] += 1
            a.adam, a.param_offset = 1, self.param_offset
            a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.lr, (a.beta1, a.beta2), a.eps, a.step = adam["lr"], adam["betas"], adam["eps"], st["step"]
        stream = c.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        self._lib.check(L.lp_exchange_step(c.byref(a), stream))

    def all_reduce(self):
        """Sum over the ranks, in place, on torch's current stream."""
        L, c = self._lib.lib(), self._ctypes
        stream = c.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        if self.fused and self.use_fused and self.one_launch:
            return self.exchange_step()
        self.handle.barrier(channel=0)                       # every rank's backward has finished
        if self.fused and self.use_fused:
            mc = c.c_void_p(self.multicast_ptr) if self.mode == "multimem" else None
            self._lib.check(L.lp_allreduce_unpack(mc, c.c_void_p(self.handle.buffer_ptrs_dev), self.accum_offset, 0,
                                                  self.texels, self.channels, self.rank, self.world, stream))
        elif self.mode == "multimem":
            self._lib.check(L.lp_allreduce_multimem(c.c_void_p(self.multicast_ptr), self.padded, self.rank, self.world, stream))
        else:
            ptrs = c.c_void_p(self.handle.buffer_ptrs_dev)
            self._lib.check(L.lp_allreduce_p2p(ptrs, self.padded, self.rank, self.world, 0, stream))
            self.handle.barrier(channel=1)                   # all slices reduced
            self._lib.check(L.lp_allreduce_p2p(ptrs, self.padded, self.rank, self.world, 1, stream))
        self.handle.barrier(channel=2)                       # every rank holds the full sum; peers are done reading
