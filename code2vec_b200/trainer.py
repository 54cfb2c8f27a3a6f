"""Train-step driver over one PathAttentionEngine per process (one process per GPU).

Single GPU: one C-ABI call per batch (c2v_train_batch_host == the reference's
``sess.run([optimizer, train_loss])``, tensorflow_model.py:80).

Data parallel (BASELINE config 4; the reference has no multi-GPU path): the
batch is sharded across ranks, tables are replicated, and the only collective is one all-reduce
(mean) of the five gradient tensors between c2v_train_step and c2v_adam_step, so every replica
applies the identical Adam update.  torch.distributed (NCCL over NVLink/NVSwitch, gloo in the CPU
tests of the host logic) is plumbing; all arithmetic stays in the engine's kernels.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import PARAM_NAMES, PathAttentionEngine

ADAM_DEFAULTS = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)   # tf.compat.v1.train.AdamOptimizer()


def allreduce_mean_(tensors, group=None):
    """In-place mean over ranks of each tensor in `tensors` (list).  NCCL averages in the
    collective itself; gloo (CPU tests) sums and scales."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    backend = dist.get_backend(group)
    handles = []
    for t in tensors:
        if backend == "nccl":
            handles.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True))
        else:
            handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    if backend != "nccl":
        for t in tensors:
            t.mul_(1.0 / world)


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n examples owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Trainer:
    def __init__(self, engine: PathAttentionEngine, keep_prob: float = 0.75, seed: int = 0, group=None,
                 adam: Optional[dict] = None):
        self.e = engine
        self.keep = float(keep_prob)
        self.seed = int(seed)
        self.group = group
        self.adam = dict(ADAM_DEFAULTS, **(adam or {}))
        torch = engine.torch
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        B, C = engine.dims.max_batch, engine.dims.max_contexts
        # device staging for the multi-GPU host path (single GPU stages inside the C ABI)
        self._dev = None
        if self.world > 1:
            i32, f32 = torch.int32, torch.float32
            self._dev = dict(src=torch.empty((B, C), dtype=i32, device=engine.dev),
                             path=torch.empty((B, C), dtype=i32, device=engine.dev),
                             tgt=torch.empty((B, C), dtype=i32, device=engine.dev),
                             mask=torch.empty((B, C), dtype=f32, device=engine.dev),
                             target=torch.empty((B,), dtype=i32, device=engine.dev))
        self._loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    # ---- inputs already resident on the device ----------------------------------------------
    def step_device(self, src, path, tgt, mask, target):
        """Forward+backward, (all-reduce), Adam.  Returns the device loss tensor (no sync)."""
        e = self.e
        t = e.adam_t + 1
        # dropout stream position: (seed, t) on every rank, offset by rank so replicas differ
        loss = e.train_step(src, path, tgt, mask, target, keep=self.keep, seed=self.seed + self.rank, step=t)
        if self.world > 1:
            allreduce_mean_([e.grads[k] for k in PARAM_NAMES], self.group)
        e.adam_step(t=t, **self.adam)
        return loss

    # ---- inputs in host memory (what train() does per batch) ---------------------------------
    def step_host(self, src, path, tgt, mask, target) -> float:
        """One training step on HOST buffers (numpy / pinned tensors); returns the loss (synchronises)."""
        e = self.e
        if self.world == 1:
            return e.train_batch_host(src, path, tgt, mask, target, keep=self.keep, seed=self.seed, **self.adam)
        torch = e.torch
        B = int(src.shape[0])
        d = self._dev
        for name, arr in (("src", src), ("path", path), ("tgt", tgt), ("mask", mask), ("target", target)):
            t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
            d[name][:B].copy_(t, non_blocking=True)
        loss = self.step_device(d["src"][:B], d["path"][:B], d["tgt"][:B], d["mask"][:B], d["target"][:B])
        self._loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream(e.dev).synchronize()
        return float(self._loss_host[0])
