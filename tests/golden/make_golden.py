"""Generates tests/golden/path_attention_golden.npz from the CPU oracle (fixed seeds).

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path and TensorFlow cannot run here (DESIGN.md
section 2), so these fixtures pin the *oracle's* outputs: `tests/test_oracle_golden.py` fails if the
oracle drifts, and the GPU tests compare the CUDA path with the same committed numbers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import path_attention_oracle as O   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "path_attention_golden.npz")
DIMS = dict(token_vocab=211, path_vocab=97, target_vocab=157, embed_dim=16, code_dim=48, max_contexts=9)
B = 12


def build():
    dims = O.Dims(**DIMS)
    params = O.init_params(dims, seed=4321)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, B, seed=1234)
    src[0, 0] = tgt[0, 0] = src[1, 0] = 3
    out = {"src": src, "pth": pth, "tgt": tgt, "mask": mask, "target": target}
    out.update({"param_" + k: v for k, v in params.items()})
    # evaluation graph
    idx, val, v, alpha, scores = O.evaluate_topk(params, src, pth, tgt, mask, k=10, normalize=False)
    out.update(code_vectors=v, attention=alpha, topk_idx=idx, topk_val=val, topk_softmax=O.softmax_over_k(val))
    # training graph, no dropout
    loss, grads, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
    out["loss"] = np.float32(loss)
    out.update({"grad_" + k: g for k, g in grads.items()})
    # training graph with the Philox dropout mask (seed, step) = (2024, 5)
    dm = O.dropout_keep_mask(seed=2024, step=5, n_rows=B * dims.max_contexts, ctx_dim=dims.ctx_dim, keep=0.75)
    loss_d, grads_d, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target, keep=0.75, dropout_mask=dm)
    out["dropout_mask_rowsum"] = dm.sum(axis=1).astype(np.int32)
    out["loss_dropout"] = np.float32(loss_d)
    out.update({"grad_dropout_" + k: g for k, g in grads_d.items()})
    # two Adam steps
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(x) for k, x in p.items()}
    vv = {k: np.zeros_like(x) for k, x in p.items()}
    for t in (1, 2):
        _, g, _ = O.train_loss_and_grads(p, src, pth, tgt, mask, target)
        O.adam_step(p, g, m, vv, t)
    out.update({"adam2_" + k: x for k, x in p.items()})
    return out


if __name__ == "__main__":
    np.savez_compressed(OUT, **build())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
