"""Helpers shared by the GPU parity tests."""
import numpy as np

from oracle import path_attention_oracle as O


def make_engine(dims: O.Dims, max_batch: int, top_k: int = 10, training: bool = True, params=None, seed=4321):
    from code2vec_b200.engine import EngineDims, PathAttentionEngine
    ed = EngineDims(dims.token_vocab, dims.path_vocab, dims.target_vocab, dims.embed_dim, dims.code_dim,
                    dims.max_contexts, max_batch, top_k)
    eng = PathAttentionEngine(ed, device=0, training=training)
    if params is None:
        params = O.init_params(dims, seed=seed)
    eng.load_params(params)
    return eng, params


def dev_batch(eng, src, pth, tgt, mask, target=None):
    import torch
    out = [eng.to_device(src, torch.int32), eng.to_device(pth, torch.int32), eng.to_device(tgt, torch.int32),
           eng.to_device(mask, torch.float32)]
    if target is not None:
        out.append(eng.to_device(target, torch.int32))
    return out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
