"""Train-step driver over one PathAttentionEngine per process (one process per GPU).

Single GPU: one C-ABI call per batch (c2v_train_batch_host == the reference's
``sess.run([optimizer, train_loss])``, tensorflow_model.py:80).

Data parallel (BASELINE config 4; the reference has no multi-GPU path): the
batch is sharded across ranks and the tables are replicated.  Two interchangeable schedules, both
leaving every replica with bit-identical parameters:

  "allreduce" : all-reduce(mean) of the five gradient tensors, then the full Adam on every rank.
  "sharded"   : (default) the model lives in one flat buffer split into two buckets -- A = the
                target table, whose gradient is complete right after the dY GEMM, B = the rest.
                Each bucket is reduce-scattered (mean) so that rank r owns 1/world of it, Adam runs
                on the owned slices only (c2v_adam_step_range: 1/world of the 9.2 GB optimizer
                traffic), and the updated slices are all-gathered back.  Bucket A's
                reduce-scatter is issued on a side stream as soon as the engine's
                "target_grads_ready" event fires, so it overlaps the context backward pass.

  "table_sharded" : (default on 2/4/8 GPUs) as "sharded" for the target table, but the two embedding
                tables are not replicated at all: they are row-sharded over the ranks and reached
                through peer memory (PathAttentionEngine.enable_table_sharding).  The forward gather
                loads rows from the owning GPU over NVLink and the backward scatter-add issues
                red.global.add to it, so no embedding gradient is ever reduced or gathered (1.13 GB
                of the 1.53 GB per step disappears from the collectives) and each rank's Adam
                touches 1/world of every table.  Ordering comes from the collectives that remain:
                the all-reduce of the TRANSFORM/ATTENTION gradients is issued after the local
                scatter, so its completion means every rank's scatter has landed (Adam may run);
                the all-gather of the updated target table is issued after the local Adam, so its
                completion means every shard is updated (the next gather may run).

  "fully_sharded" : (BASELINE config 5) nothing big is replicated.  Embedding tables as in
                "table_sharded"; the target table is row-sharded in contiguous blocks and each rank's
                engine is built for its LOCAL target rows and the GLOBAL batch (make_fully_sharded_engine).
                A step is phase-split (c2v_context_forward / c2v_target_forward / c2v_lse_combine /
                c2v_target_backward / c2v_context_backward) and moves only small tensors between the
                phases: all-gather of code vectors [Bt, D] and targets, all-gather of the per-row
                (max, sum exp) partials and all-reduce of the true logits [Bt], reduce-scatter of dv
                [Bt, D].  No logits slab and no table gradient ever crosses NVLink; each rank's Adam
                covers 1/world of all three tables.
                In the tensor-core modes the per-row partials are (c, sum exp(s - c)) of the exp_slab
                schedule (c = the true logit on the rank that owns the class, 0 elsewhere) -- the same
                combine formula -- and the softmax's normalisation is applied as per-row factors inside
                the two gradient GEMMs.  Trainer(allow_single_rank=True) keeps
                this schedule in a process group of ONE rank, so a one-GPU box can run the whole path
                (tests/test_gpu_dp.py).

torch.distributed (NCCL over NVLink/NVSwitch; gloo in the CPU tests of the host logic) is plumbing;
all arithmetic stays in the engine's kernels.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import PARAM_NAMES, PathAttentionEngine

ADAM_DEFAULTS = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)   # tf.compat.v1.train.AdamOptimizer()


def _dist():
    import torch.distributed as dist
    return dist


def allreduce_mean_(tensors, group=None):
    """In-place mean over ranks of each tensor in `tensors` (list).  NCCL averages in the
    collective itself; gloo (CPU tests) sums and scales."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    backend = dist.get_backend(group)
    op = dist.ReduceOp.AVG if backend == "nccl" else dist.ReduceOp.SUM
    handles = [dist.all_reduce(t, op=op, group=group, async_op=True) for t in tensors]
    for h in handles:
        h.wait()
    if backend != "nccl":
        for t in tensors:
            t.mul_(1.0 / world)


def reduce_scatter_mean(out_shard, flat, group=None, async_op=False):
    """out_shard <- mean over ranks of this rank's 1/world slice of `flat` (len(flat) % world == 0)."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == "nccl":
        return dist.reduce_scatter_tensor(out_shard, flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    # gloo has no reduce-scatter: all-reduce a copy and keep the owned slice (CPU tests only)
    tmp = flat.clone()
    dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
    n = flat.numel() // world
    out_shard.copy_(tmp[rank * n:(rank + 1) * n] / world)
    return None


def all_gather_flat(flat, shard, group=None, async_op=False):
    """flat <- concatenation over ranks of `shard` (len(flat) == world * len(shard))."""
    dist = _dist()
    if dist.get_backend(group) == "nccl":
        return dist.all_gather_into_tensor(flat, shard, group=group, async_op=async_op)
    world = dist.get_world_size(group)
    n = shard.numel()
    parts = [flat[i * n:(i + 1) * n] for i in range(world)]
    dist.all_gather(parts, shard.clone(), group=group)
    return None


def target_row_block(n_targets: int, rank: int, world: int):
    """Contiguous block [row0, row1) of target-table rows owned by `rank` in the fully sharded schedule."""
    per = (n_targets + world - 1) // world
    row0 = min(rank * per, n_targets)
    return row0, min(row0 + per, n_targets)


def make_fully_sharded_engine(dims, local_batch: int, device: int, group=None, training: bool = True):
    """Engine for the fully sharded schedule: `dims` are the GLOBAL model dims (EngineDims); the returned
    engine holds this rank's block of target rows and is sized for the global batch."""
    from dataclasses import replace
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    row0, row1 = target_row_block(dims.target_vocab, rank, world)
    local = replace(dims, target_vocab=max(row1 - row0, 1), max_batch=local_batch * world)
    eng = PathAttentionEngine(local, device=device, training=training)
    eng.global_target_vocab, eng.target_row0, eng.local_batch = dims.target_vocab, row0, local_batch
    return eng


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Trainer:
    def __init__(self, engine: PathAttentionEngine, keep_prob: float = 0.75, seed: int = 0, group=None,
                 adam: Optional[dict] = None, schedule: str = "table_sharded", lazy_adam: bool = True,
                 fuse_target_adam: bool = True, push_grads: bool = False, allow_single_rank: bool = False):
        self.e = engine
        self.keep = float(keep_prob)
        self.seed = int(seed)
        self.group = group
        self.adam = dict(ADAM_DEFAULTS, **(adam or {}))
        torch = engine.torch
        dist = _dist()
        grouped = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if grouped else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        # allow_single_rank: keep the fully sharded schedule (IPC-homed tables, phase-split entry points, collectives) in a
        # process group of ONE rank -- no use in production, but it lets a one-GPU box run that whole code path
        self.multi = self.world > 1 or (allow_single_rank and grouped and schedule == "fully_sharded")
        self.schedule = schedule if self.multi else "single"
        if self.schedule in ("table_sharded", "fully_sharded") and self.world not in (1, 2, 4, 8):
            self.schedule = "sharded"
        if self.schedule == "fully_sharded":
            if not hasattr(engine, "target_row0"):
                raise ValueError("the fully_sharded schedule needs an engine from make_fully_sharded_engine()")
            with torch.cuda.device(engine.dev):
                engine.enable_table_sharding(group, push_grads=push_grads)
            engine.set_option("grad_scale_inverse", 1)     # dv already carries the 1/global-batch factor
            Bl, Bt, D = engine.local_batch, engine.local_batch * self.world, engine.dims.code_dim
            f32, i32, dev = torch.float32, torch.int32, engine.dev
            z = lambda shape, dt=f32: torch.zeros(shape, dtype=dt, device=dev)
            self._fs = dict(v_local=z((Bl, D)), v_all=z((Bt, D)), tgt_all=z((Bt,), i32), rmax=z((Bt,)), rsum=z((Bt,)),
                            tlogit=z((Bt,)), maxes=z((self.world, Bt)), sums=z((self.world, Bt)), lse=z((Bt,)),
                            dv_part=z((Bt, D)), dv_local=z((Bl, D)), loss=z((1,)), token=z((1,)))
            layout, total = engine.flat_layout()
            small0 = [off for k, off, n in layout if k == "W"][0]
            self._small = (small0, total)
        B, C = engine.dims.max_batch, engine.dims.max_contexts
        self._dev = None
        if self.multi:
            i32, f32 = torch.int32, torch.float32
            self._dev = dict(src=torch.empty((B, C), dtype=i32, device=engine.dev),
                             path=torch.empty((B, C), dtype=i32, device=engine.dev),
                             tgt=torch.empty((B, C), dtype=i32, device=engine.dev),
                             mask=torch.empty((B, C), dtype=f32, device=engine.dev),
                             target=torch.empty((B,), dtype=i32, device=engine.dev))
        if self.schedule == "table_sharded":
            with torch.cuda.device(engine.dev):
                engine.enable_table_sharding(group, push_grads=push_grads)
            (a0, a1), _ = engine.bucket_bounds()
            layout, total = engine.flat_layout()
            small0 = [off for k, off, n in layout if k == "W"][0]         # W, a: the tail of the flat buffer
            w, r = self.world, self.rank
            na = (a1 - a0) // w
            self._bucket = [(a0, a1, a0 + r * na, a0 + (r + 1) * na)]
            self._small = (small0, total)
            self._gshard = [torch.empty(na, dtype=torch.float32, device=engine.dev)]
            self._side = torch.cuda.Stream(device=engine.dev)
            self._ev_tgt = torch.cuda.Event()
            with torch.cuda.device(engine.dev):
                engine.set_event("target_grads_ready", self._ev_tgt)
        if self.schedule in ("sharded", "table_sharded"):
            engine.set_option("dy_late", 0)     # these schedules start the target table's reduce-scatter right after dY
        if self.schedule == "sharded":
            (a0, a1), (b0, b1) = engine.bucket_bounds()
            w, r = self.world, self.rank
            assert (a1 - a0) % w == 0 and (b1 - b0) % w == 0
            na, nb = (a1 - a0) // w, (b1 - b0) // w
            self._bucket = [(a0, a1, a0 + r * na, a0 + (r + 1) * na), (b0, b1, b0 + r * nb, b0 + (r + 1) * nb)]
            self._gshard = [torch.empty(na, dtype=torch.float32, device=engine.dev),
                            torch.empty(nb, dtype=torch.float32, device=engine.dev)]
            self._side = torch.cuda.Stream(device=engine.dev)
            self._ev_tgt = torch.cuda.Event()
            with torch.cuda.device(engine.dev):
                engine.set_event("target_grads_ready", self._ev_tgt)
        if self.schedule == "single" and lazy_adam and engine.training:
            engine.set_option("lazy_adam", 1)       # exact, see c2v_b200.h; the multi-GPU schedules stay dense
        # target-table Adam inside the dY epilogue (bit-identical, see c2v_arm_target_adam): wherever the
        # target gradient is complete on this rank without a collective
        self.fuse_tgt = bool(fuse_target_adam) and self.schedule in ("single", "fully_sharded") and engine.training
        if self.schedule == "single":
            engine.set_option("fuse_target_adam", 1 if self.fuse_tgt else 0)     # for train_batch_host
            # dY (+ the target table's Adam step) straight after dv: on one GPU it is HBM-bound like the scatter-add
            # it would otherwise share the memory system with (dy_late 0/1/2 all measure within 1 %)
            engine.set_option("dy_late", 0)
        self._loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    # ---- inputs already resident on the device ----------------------------------------------
    def step_device(self, src, path, tgt, mask, target, next_batch=None):
        """Forward+backward, gradient exchange, Adam.  Returns the device loss tensor (no sync).
        next_batch: optional (src, path, tgt) device tensors of the batch the NEXT call will get -- with lazy
        Adam the deferred updates of its rows then overlap this step's backward GEMMs (c2v_hint_next_batch)."""
        if self.schedule == "fully_sharded":
            return self._fully_sharded_step(src, path, tgt, mask, target)
        e = self.e
        t = e.adam_t + 1
        if self.fuse_tgt:
            e.arm_target_adam(t, **self.adam)
            if next_batch is not None and self.schedule == "single":
                e.hint_next_batch(*next_batch[:3])
        # dropout stream position (seed, t); replicas use different seeds so their masks differ
        loss = e.train_step(src, path, tgt, mask, target, keep=self.keep, seed=self.seed + self.rank, step=t)
        if self.schedule == "single":
            e.adam_step(t=t, **self.adam)
        elif self.schedule == "allreduce":
            allreduce_mean_([e.grads[k] for k in PARAM_NAMES], self.group)
            e.adam_step(t=t, **self.adam)
        elif self.schedule == "table_sharded":
            self._table_sharded_update(t)
        elif self.schedule == "fully_sharded":
            raise RuntimeError("fully_sharded uses step_device_fully_sharded (phase-split step)")
        else:
            self._sharded_update(t)
        return loss

    def step_ring(self, ring, rows: int, loss_out):
        """One training step on the next filled slot of a PinnedBatchRing, fully asynchronous (single GPU)."""
        if self.schedule != "single":
            raise RuntimeError("step_ring is single-GPU")
        ring.submit_next(self.e, rows, loss_out, keep=self.keep, seed=self.seed, **self.adam)

    def step_device_sampled(self, src, path, tgt, mask, target, sampled, logq_true, logq_sampled):
        """BASELINE config 3: one training step with the sampled softmax (c2v_sampled_train_step) + Adam.  Single
        GPU.  With lazy Adam the target table's rows are updated lazily too (only the B + S rows the step reads)."""
        if self.schedule != "single":
            raise RuntimeError("the sampled-softmax step is single-GPU")
        e = self.e
        t = e.adam_t + 1
        loss = e.sampled_train_step(src, path, tgt, mask, target, sampled, logq_true, logq_sampled, keep=self.keep,
                                    seed=self.seed, step=t)
        e.adam_step(t=t, **self.adam)
        return loss

    def _fully_sharded_step(self, src, path, tgt, mask, target):
        e, dist, fs = self.e, _dist(), self._fs
        t = e.adam_t + 1
        B, W = int(src.shape[0]), self.world
        if B != e.local_batch:
            raise ValueError("fully_sharded needs the same local batch (%d) on every rank" % e.local_batch)
        seed = self.seed + self.rank
        e.context_forward(src, path, tgt, mask, fs["v_local"], keep=self.keep, seed=seed, step=t)
        dist.all_gather_into_tensor(fs["v_all"], fs["v_local"], group=self.group)
        dist.all_gather_into_tensor(fs["tgt_all"], target, group=self.group)
        e.target_forward(fs["v_all"], fs["tgt_all"], e.target_row0, fs["rmax"], fs["rsum"], fs["tlogit"])
        dist.all_gather_into_tensor(fs["maxes"].view(-1), fs["rmax"], group=self.group)
        dist.all_gather_into_tensor(fs["sums"].view(-1), fs["rsum"], group=self.group)
        dist.all_reduce(fs["tlogit"], op=dist.ReduceOp.SUM, group=self.group)
        e.lse_combine(fs["maxes"], fs["sums"], fs["tlogit"], fs["lse"], fs["loss"])
        if self.fuse_tgt:
            e.arm_target_adam(t, **self.adam)
        e.target_backward(fs["v_all"], fs["lse"], fs["tgt_all"], e.target_row0, fs["dv_part"])
        dist.reduce_scatter_tensor(fs["dv_local"], fs["dv_part"], op=dist.ReduceOp.SUM, group=self.group)
        e.context_backward(src, path, tgt, mask, fs["dv_local"], keep=self.keep, seed=seed, step=t)
        s0, s1 = self._small
        # sum (not mean): dv already carries 1/global batch.  Completion == every rank's scatter-add has landed.
        dist.all_reduce(e.flat_grads[s0:s1], op=dist.ReduceOp.SUM, group=self.group)
        if getattr(e, "push_grads", False):
            e.apply_scatter_inbox()          # every peer's rows have landed in this rank's inbox: fold them into the shards
        for name in ("tok", "path"):
            e.adam_step_range(e.shard_params[name], e.shard_grads[name], e.shard_m[name], e.shard_v[name], t,
                              zero_grad=True, **self.adam)
        if e.get_option("target_adam_fused_step") == t:
            e.set_option("target_adam_fused_step", 0)       # the dY epilogue already applied this rank's target rows
        else:
            e.adam_step_range(e.params["tgt"], e.grads["tgt"], e.adam_m["tgt"], e.adam_v["tgt"], t, **self.adam)
        e.adam_step_range(e.flat_params[s0:s1], e.flat_grads[s0:s1], e.flat_m[s0:s1], e.flat_v[s0:s1], t, **self.adam)
        # every shard must be updated before any rank's next gather reads it
        dist.all_reduce(fs["token"], group=self.group)
        return fs["loss"]

    def _table_sharded_update(self, t: int):
        e, torch = self.e, self.e.torch
        dist = _dist()
        main = torch.cuda.current_stream(e.dev)
        (a0, a1, alo, ahi), = self._bucket
        s0, s1 = self._small
        self._side.wait_event(self._ev_tgt)
        with torch.cuda.stream(self._side):
            wa = reduce_scatter_mean(self._gshard[0], e.flat_grads[a0:a1], self.group, async_op=True)
        # TRANSFORM / ATTENTION gradients (0.6 MB): mean over ranks; issued after the local scatter-add, so
        # its completion also tells this rank that every peer's red.adds into its shards have landed
        ws = dist.all_reduce(e.flat_grads[s0:s1], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        ws.wait()
        if getattr(e, "push_grads", False):
            e.apply_scatter_inbox()
        for name in ("tok", "path"):
            e.adam_step_range(e.shard_params[name], e.shard_grads[name], e.shard_m[name], e.shard_v[name], t,
                              zero_grad=True, **self.adam)
        e.adam_step_range(e.flat_params[s0:s1], e.flat_grads[s0:s1], e.flat_m[s0:s1], e.flat_v[s0:s1], t, **self.adam)
        if wa is not None:
            wa.wait()
        main.wait_stream(self._side)
        e.adam_step_range(e.flat_params[alo:ahi], self._gshard[0], e.flat_m[alo:ahi], e.flat_v[alo:ahi], t, **self.adam)
        # issued after every local Adam launch: completion == all ranks' shards are updated
        ga = all_gather_flat(e.flat_params[a0:a1], e.flat_params[alo:ahi], self.group, async_op=True)
        if ga is not None:
            ga.wait()

    def _sharded_update(self, t: int):
        e, torch = self.e, self.e.torch
        main = torch.cuda.current_stream(e.dev)
        (a0, a1, alo, ahi), (b0, b1, blo, bhi) = self._bucket
        # bucket A: may start as soon as dY is done (event recorded inside c2v_train_step)
        self._side.wait_event(self._ev_tgt)
        with torch.cuda.stream(self._side):
            wa = reduce_scatter_mean(self._gshard[0], e.flat_grads[a0:a1], self.group, async_op=True)
        wb = reduce_scatter_mean(self._gshard[1], e.flat_grads[b0:b1], self.group, async_op=True)
        for wk in (wa, wb):
            if wk is not None:
                wk.wait()                       # current (main) stream waits for the collective
        main.wait_stream(self._side)
        e.adam_step_range(e.flat_params[alo:ahi], self._gshard[0], e.flat_m[alo:ahi], e.flat_v[alo:ahi], t, **self.adam)
        e.adam_step_range(e.flat_params[blo:bhi], self._gshard[1], e.flat_m[blo:bhi], e.flat_v[blo:bhi], t, **self.adam)
        ga = all_gather_flat(e.flat_params[a0:a1], e.flat_params[alo:ahi], self.group, async_op=True)
        gb = all_gather_flat(e.flat_params[b0:b1], e.flat_params[blo:bhi], self.group, async_op=True)
        for wk in (ga, gb):
            if wk is not None:
                wk.wait()

    # ---- inputs in host memory (what train() does per batch) ---------------------------------
    def step_host(self, src, path, tgt, mask, target, next_batch=None) -> float:
        """One training step on HOST buffers (numpy / pinned tensors); returns the loss (synchronises).
        next_batch: optional host (src, path, tgt) of the following batch (see step_device)."""
        e = self.e
        if not self.multi:
            if next_batch is not None and self.fuse_tgt:
                e.hint_next_batch_host(*next_batch[:3])
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
