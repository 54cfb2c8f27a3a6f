"""Synthetic path-context bags of a given shape for benchmarks and demos.

Valid slots get indices in [1, V); slots past the bag's length are index 0 in all three parts with
mask 0 -- exactly what the reader emits for padding (reference path_context_reader.py:79-83,
210-214).  The same generator exists in the test oracle (oracle/path_attention_oracle.py) so tests
and bench draw identical batches; tests/test_host_surface.py checks the two agree."""
from __future__ import annotations

import numpy as np


def synthetic_batch(token_vocab: int, path_vocab: int, target_vocab: int, max_contexts: int, batch: int, seed: int = 1234,
                    full_bags: bool = False, zipf: bool = False, normal_bags: bool = False):
    rng = np.random.default_rng(seed)
    B, C = batch, max_contexts

    def draw(hi, shape):
        if zipf:
            r = rng.zipf(1.2, size=shape)
            return (1 + (r - 1) % (hi - 1)).astype(np.int32)
        return rng.integers(1, hi, size=shape, dtype=np.int32)

    src = draw(token_vocab, (B, C))
    pth = draw(path_vocab, (B, C))
    tgt = draw(token_vocab, (B, C))
    if full_bags:
        n_valid = np.full(B, C)
    elif normal_bags:                      # SURVEY 8d second run: n_b ~ clip(N(0.6 C, 0.3 C), 1, C)
        n_valid = np.clip(np.rint(rng.normal(0.6 * C, 0.3 * C, size=B)), 1, C).astype(np.int64)
    else:
        n_valid = rng.integers(1, C + 1, size=B)
    valid = np.arange(C)[None, :] < n_valid[:, None]
    src = np.where(valid, src, 0).astype(np.int32)
    pth = np.where(valid, pth, 0).astype(np.int32)
    tgt = np.where(valid, tgt, 0).astype(np.int32)
    mask = valid.astype(np.float32)
    target = rng.integers(1, target_vocab, size=B, dtype=np.int32)
    return src, pth, tgt, mask, target
