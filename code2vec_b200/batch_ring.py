"""Pinned-host batch ring + copy stream: the asynchronous half of the batcher that replaces
`path_context_reader.py:119-151` (tf.data map -> batch -> prefetch(40)).

The reader thread writes every training batch straight into one of `slots` page-locked host buffers
(PinnedBatchRing.acquire); the training loop uploads the next filled slot on an engine-owned copy stream into one
of two device buffers and launches the step on the compute stream behind an event (upload_next), so the
host -> device copy of batch t+1 overlaps the kernels of batch t and the host never blocks on the GPU inside
the loop.  torch supplies pinned memory, streams and events (plumbing only): no arithmetic happens here.
"""
from __future__ import annotations

import threading
from typing import List, Optional

import numpy as np

_FIELDS = (("src", "int32"), ("path", "int32"), ("tgt", "int32"), ("mask", "float32"))


class _Slot:
    def __init__(self, torch, batch: int, contexts: int):
        self.t = {}
        for name, dt in _FIELDS:
            self.t[name] = torch.empty((batch, contexts), dtype=getattr(torch, dt)).pin_memory()
        self.t["target"] = torch.empty((batch,), dtype=torch.int32).pin_memory()
        # numpy views of the same page-locked memory: what the tensoriser / shuffle pool writes into
        self.np = {k: v.numpy() for k, v in self.t.items()}
        self.free = threading.Event()
        self.free.set()
        self.h2d_done = None             # torch.cuda.Event of the last upload out of this slot

    def arrays(self):
        """The five row arrays in the reader's column order (src, path, tgt, mask, target)."""
        return (self.np["src"], self.np["path"], self.np["tgt"], self.np["mask"], self.np["target"])


class PinnedBatchRing:
    def __init__(self, torch, device, batch: int, contexts: int, slots: int = 10):
        if slots < 3:
            raise ValueError("the ring needs at least 3 slots")
        self.torch, self.device = torch, device
        self.batch, self.contexts = int(batch), int(contexts)
        self.slots: List[_Slot] = [_Slot(torch, batch, contexts) for _ in range(slots)]
        with torch.cuda.device(device):
            self.dev = [{k: torch.empty_like(v, device=device) for k, v in self.slots[0].t.items()} for _ in range(2)]
            self.copy_stream = torch.cuda.Stream(device=device)
        self.compute_done: List[Optional[object]] = [None, None]
        self._fill = 0                    # producer cursor (reader thread)
        self._use = 0                     # consumer cursor (training loop)
        self.closed = False

    # ---- producer side (reader thread) ---------------------------------------------------------
    def acquire(self) -> _Slot:
        """The next slot in FIFO order, once its previous contents have left for the device."""
        s = self.slots[self._fill % len(self.slots)]
        while not s.free.wait(timeout=0.1):
            if self.closed:
                raise RuntimeError("batch ring closed")
        if s.h2d_done is not None:
            s.h2d_done.synchronize()
            s.h2d_done = None
        s.free.clear()
        self._fill += 1
        return s

    # ---- consumer side (training loop) ------------------------------------------------------------
    def upload_next(self, rows: int):
        """Queue the host -> device copy of the next filled slot (first `rows` rows) on the copy stream and make the
        current (compute) stream wait for it.  Returns (device tensors dict, device buffer index)."""
        torch = self.torch
        s = self.slots[self._use % len(self.slots)]
        i = self._use % 2
        self._use += 1
        d = self.dev[i]
        with torch.cuda.stream(self.copy_stream):
            if self.compute_done[i] is not None:
                self.copy_stream.wait_event(self.compute_done[i])      # the step that last read this device buffer
            for k in ("src", "path", "tgt", "mask", "target"):
                d[k][:rows].copy_(s.t[k][:rows], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        s.h2d_done = ev
        s.free.set()                      # the producer synchronises on h2d_done before it refills the slot
        torch.cuda.current_stream(self.device).wait_event(ev)
        return {k: v[:rows] for k, v in d.items()}, i

    def submit_next(self, engine, rows: int, loss_out, **step_args):
        """The whole step of the next filled slot as ONE C-ABI call (c2v_train_batch_async): upload on the engine's copy
        stream, the step queued behind it, nothing waited for.  The Python side of a step is then a single call that
        releases the GIL -- the reader thread, which is what bounds a text-fed run, keeps the interpreter."""
        s = self.slots[self._use % len(self.slots)]
        self._use += 1
        ev = self.torch.cuda.Event()
        t = s.t
        engine.train_batch_async(t["src"], t["path"], t["tgt"], t["mask"], t["target"], rows, loss_out, upload_done=ev, **step_args)
        s.h2d_done = ev
        s.free.set()

    def mark_compute_done(self, i: int):
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream(self.device))
        self.compute_done[i] = ev

    @property
    def bytes_per_batch(self) -> int:
        return sum(int(v.numel()) * 4 for v in self.slots[0].t.values())

    def close(self):
        self.closed = True
        for s in self.slots:
            s.free.set()
