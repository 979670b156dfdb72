"""Optimiser + data-parallel gradient exchange of the training step.

Reference: `configure_optimizers` (sgm/models/diffusion.py:310-373) hands the trainable ('pose')
parameters to torch.optim.AdamW; Lightning's DDP strategy all-reduces (mean) their gradients.

Here the trainable parameters (12 pose blocks, ~66.7 M values at SDXL size) are re-homed into ONE
flat fp32 buffer, their `.grad`s into a second one (the parameters / grads become views, names and
shapes unchanged, so checkpoints are unaffected):
  * the optimiser update is one fused AdamW launch over the flat buffer (cd360_adamw_step);
  * the gradient exchange is one all-reduce per BUCKET of the flat gradient buffer (buckets =
    consecutive pose blocks in backward order), issued on a side stream so that early buckets
    overlap the rest of the backward when `reduce_bucket` is called as blocks finish, or all at
    once from `step()`.  The division by the world size is folded into AdamW's `grad_scale`.
NCCL over NVLink on the B200 box; gloo in the CPU tests of the bucket logic.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import ops


def _align(n: int, a: int = 4) -> int:
    return (n + a - 1) // a * a


class FlatParams:
    """Flat fp32 storage for a list of named parameters; every slice starts 16-byte aligned (the
    weight-gradient GEMMs write into the gradient views directly)."""

    def __init__(self, named: Sequence[Tuple[str, torch.nn.Parameter]]):
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += _align(p.numel())
        self.numel = off
        self.data = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(off, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.data[o:o + n].copy_(p.detach().float().reshape(-1))
            p.data = self.data[o:o + n].view(p.shape)
            p.grad = self.grad[o:o + n].view(p.shape)

    def buckets_by_prefix(self, depth_marker: str = ".pose") -> List[Tuple[int, int]]:
        """[start, end) ranges of the flat buffers, one per pose block (parameters are grouped by
        the module path before `.pose…`), in parameter order."""
        out, cur, start = [], None, 0
        for name, o in zip(self.names, self.offsets):
            key = name.split(depth_marker)[0]
            if cur is None:
                cur = key
            elif key != cur:
                out.append((start, o))
                start, cur = o, key
        out.append((start, self.numel))
        return out


class PoseAdamW:
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) on `FlatParams`."""

    def __init__(self, named, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, group=None,
                 **unused):
        self.flat = FlatParams(named)
        self.lr, self.betas, self.eps, self.weight_decay = lr, tuple(betas), eps, weight_decay
        self.m = torch.zeros_like(self.flat.data)
        self.v = torch.zeros_like(self.flat.data)
        self.steps = 0
        self.group = group
        self.buckets = self.flat.buckets_by_prefix()
        self._pending: list = []
        self._comm_stream = None
        # True while the step runs as a CUDA-graph replay (GraphedTrainStep): the per-block callbacks of
        # the backward walk do nothing and `step()` all-reduces every bucket after the replay
        self.suspend_overlap = False
        self.on_step = None
        self.param_groups = [{"lr": lr, "params": self.flat.params}]

    # ---- data-parallel gradient exchange ----------------------------------------------------------
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce_bucket(self, i: int, force: bool = False):
        """Start the (sum) all-reduce of bucket i of the flat gradient buffer.  On CUDA the
        collective runs on a side stream ordered after the kernels already queued on the current
        stream, so it overlaps whatever the backward launches next."""
        if self._world() == 1 or (self.suspend_overlap and not force):
            return
        lo, hi = self.buckets[i]
        g = self.flat.grad[lo:hi]
        if g.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=g.device)
            self._comm_stream.wait_stream(torch.cuda.current_stream(g.device))
            with torch.cuda.stream(self._comm_stream):
                self._pending.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._pending.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def attach_overlap(self, unet):
        """Start each pose block's bucket all-reduce the moment the backward walk has written that
        block's gradients (blocks finish in reverse forward order, so the exchange of the decoder's
        blocks overlaps the backward of everything upstream)."""
        prefixes = []
        for name in self.flat.names:
            key = name.split(".pose")[0]
            if not prefixes or prefixes[-1] != key:
                prefixes.append(key)
        assert len(prefixes) == len(self.buckets)
        index = {k: i for i, k in enumerate(prefixes)}
        for name, block in unet.pose_blocks():
            i = index[name]
            block.__dict__["_grads_ready"] = (lambda i=i: self.reduce_bucket(i))

    def reduce_all(self):
        for i in range(len(self.buckets)):
            self.reduce_bucket(i, force=True)

    def wait_reduce(self):
        """Make the current stream wait for every exchange started since the last call.  (Also valid
        during CUDA-graph capture: the waits join the communication stream back into the capture;
        with nothing pending no cross-stream dependency is created.)"""
        had = bool(self._pending)
        for w in self._pending:
            w.wait()
        self._pending = []
        if had and self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)

    # ---- optimiser -------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self.flat.grad.zero_()

    def step(self, reduce: bool = True):
        world = self._world()
        if reduce and world > 1:
            if not self._pending:
                self.reduce_all()
            self.wait_reduce()
        self.steps += 1
        ops.adamw_step(self.flat.data, self.flat.grad, self.m, self.v, lr=self.param_groups[0]["lr"],
                       beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                       weight_decay=self.weight_decay, step=self.steps, grad_scale=1.0 / world)
        if self.on_step is not None:
            self.on_step()

    def state_dict(self):
        return {"step": self.steps, "exp_avg": self.m, "exp_avg_sq": self.v, "names": self.flat.names,
                "offsets": self.flat.offsets}
