"""Lazy-but-exact Adam (option "lazy_adam"): deferring the zero-gradient updates of untouched embedding
rows must not change results.  Checked (i) against the oracle's dense TF1 Adam over several steps
with different batches, so rows are touched, left alone for a few steps, and touched again;
(ii) bit for bit against the engine's own dense Adam on rows whose gradients are free of
atomic-order effects (referenced by exactly one context)."""
import numpy as np
import pytest

from oracle import path_attention_oracle as O
from tests.util import dev_batch, make_engine

pytestmark = pytest.mark.gpu

DIMS = O.Dims(token_vocab=4001, path_vocab=2003, target_vocab=301, embed_dim=32, code_dim=96, max_contexts=10)
B, STEPS = 8, 7


def _batches():
    return [O.synthetic_batch(DIMS, B, seed=100 + s) for s in range(STEPS)]


@pytest.mark.parametrize("math_mode", [0, 1])
def test_lazy_adam_matches_dense_oracle_and_dense_engine(math_mode):
    import torch
    batches = _batches()
    lazy, params0 = make_engine(DIMS, max_batch=B)
    dense, _ = make_engine(DIMS, max_batch=B, params=params0)
    for eng in (lazy, dense):
        eng.set_option("math_mode", math_mode)
    lazy.set_option("lazy_adam", 1)
    assert lazy.get_option("lazy_adam") == 1
    for s, (src, pth, tgt, mask, target) in enumerate(batches):
        for eng in (lazy, dense):
            d = dev_batch(eng, src, pth, tgt, mask, target)
            eng.train_step(*d, keep=1.0)
            eng.adam_step()
    got_lazy, got_dense = lazy.export_params(), dense.export_params()      # export replays deferred updates
    if math_mode == 0:
        params = {k: v.copy() for k, v in params0.items()}
        m = {k: np.zeros_like(p) for k, p in params.items()}
        v = {k: np.zeros_like(p) for k, p in params.items()}
        for s, (src, pth, tgt, mask, target) in enumerate(batches):
            _, g, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
            O.adam_step(params, g, m, v, s + 1)
        for k in O.PARAM_NAMES:
            assert np.abs(got_lazy[k] - params[k]).max() < 5e-5, k
    for k in O.PARAM_NAMES:
        assert np.abs(got_lazy[k] - got_dense[k]).max() < 2e-6, k
    # rows referenced by exactly one context in exactly one step: no atomic-order freedom -> bit-identical
    for name, cols in (("tok", (0, 2)), ("path", (1,))):
        counts = np.zeros(got_lazy[name].shape[0], dtype=np.int64)
        for batch in batches:
            for c in cols:
                idx = batch[c][batch[3] > 0]
                np.add.at(counts, idx, 1)
        once = counts == 1
        assert once.sum() > 50
        assert np.array_equal(got_lazy[name][once], got_dense[name][once]), name
        assert np.abs(got_lazy[name][once] - params0[name][once]).max() > 1e-4      # they did move
    # Adam slots agree too (they are part of a checkpoint)
    assert torch.allclose(lazy.adam_m["tok"], dense.adam_m["tok"], atol=1e-7)
    assert torch.allclose(lazy.adam_v["path"], dense.adam_v["path"], atol=1e-9)


def test_lazy_adam_through_host_entry_point_and_mode_switch():
    batches = _batches()
    eng, params0 = make_engine(DIMS, max_batch=B)
    ref, _ = make_engine(DIMS, max_batch=B, params=params0)
    eng.set_option("lazy_adam", 1)
    for i, (src, pth, tgt, mask, target) in enumerate(batches):
        la = eng.train_batch_host(src, pth, tgt, mask, target, keep=1.0)
        lb = ref.train_batch_host(src, pth, tgt, mask, target, keep=1.0)
        assert abs(la - lb) < 1e-5
        if i == 3:
            eng.set_option("lazy_adam", 0)          # flush, continue dense ...
        if i == 4:
            eng.set_option("lazy_adam", 1)          # ... and back
    a, b = eng.export_params(), ref.export_params()
    for k in O.PARAM_NAMES:
        assert np.abs(a[k] - b[k]).max() < 2e-6, k
    # predictions after training read caught-up rows as well
    src, pth, tgt, mask, _ = batches[0]
    ia, va, ca, _ = eng.predict_batch_host(src, pth, tgt, mask)
    ib, vb, cb, _ = ref.predict_batch_host(src, pth, tgt, mask)
    assert np.abs(ca - cb).max() < 1e-5


@pytest.mark.parametrize("entry", ["device", "host"])
def test_next_batch_hint_runs_the_deferred_updates_early_and_changes_nothing(entry):
    """Trainer("single") = lazy Adam + target Adam in the dY epilogue + next-batch hint (the deferred row
    updates of batch t+1 run during step t's backward) against a plain dense engine: train_step + adam_step."""
    import torch
    from code2vec_b200.trainer import Trainer
    batches = _batches()
    fast, params0 = make_engine(DIMS, max_batch=B)
    dense, _ = make_engine(DIMS, max_batch=B, params=params0)
    for eng in (fast, dense):
        eng.set_option("math_mode", 1)
    tr = Trainer(fast, keep_prob=1.0, seed=3)
    assert tr.schedule == "single" and tr.fuse_tgt and fast.get_option("lazy_adam") == 1
    for s, (src, pth, tgt, mask, target) in enumerate(batches):
        nxt = batches[s + 1][:3] if s + 1 < len(batches) else None
        if entry == "device":
            d = dev_batch(fast, src, pth, tgt, mask, target)
            dn = None if nxt is None else dev_batch(fast, nxt[0], nxt[1], nxt[2], mask)[:3]
            la = float(tr.step_device(*d, next_batch=dn).cpu()[0])
        else:
            la = tr.step_host(src, pth, tgt, mask, target, next_batch=nxt)
        lb = float(dense.train_step(*dev_batch(dense, src, pth, tgt, mask, target), keep=1.0).cpu()[0])
        dense.adam_step()
        assert abs(la - lb) < 1e-5
    # the first step cannot run ahead (no hyper-parameters on record yet), the last one has no next batch
    assert fast.get_option("early_catchup_count") == STEPS - 2
    a, b = fast.export_params(), dense.export_params()
    assert np.array_equal(a["tgt"], b["tgt"])                    # epilogue Adam == adam_kernel, bit for bit
    for k in O.PARAM_NAMES:
        assert np.abs(a[k] - b[k]).max() < 2e-6, k
    for name, cols in (("tok", (0, 2)), ("path", (1,))):
        counts = np.zeros(a[name].shape[0], dtype=np.int64)
        for batch in batches:
            for c in cols:
                np.add.at(counts, batch[c][batch[3] > 0], 1)
        once = counts == 1
        assert once.sum() > 50 and np.array_equal(a[name][once], b[name][once]), name
    assert torch.allclose(fast.adam_m["tok"], dense.adam_m["tok"], atol=1e-7)
    assert torch.allclose(fast.adam_v["path"], dense.adam_v["path"], atol=1e-9)
