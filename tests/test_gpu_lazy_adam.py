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


@pytest.mark.parametrize("math_mode", [0, 1, 2])
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
    if math_mode != 1:          # the two fp32-class modes follow the oracle's trajectory
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


@pytest.mark.parametrize("period", [0, 1, 3, 32])
def test_sweep_period_changes_nothing(period):
    """Option "adam_sweep_period": each step also brings a 1/R slice of every table up to date, so no row is ever
    more than R steps behind.  The deferred steps are only applied earlier -- 40 steps (more than one full sweep
    at R = 32) with different batches must leave the model where the dense engine leaves it."""
    steps = 40
    batches = [O.synthetic_batch(DIMS, B, seed=500 + s) for s in range(steps)]
    lazy, params0 = make_engine(DIMS, max_batch=B)
    dense, _ = make_engine(DIMS, max_batch=B, params=params0)
    lazy.set_option("lazy_adam", 1)
    lazy.set_option("adam_sweep_period", period)
    assert lazy.get_option("adam_sweep_period") == period
    for src, pth, tgt, mask, target in batches:
        for eng in (lazy, dense):
            eng.train_step(*dev_batch(eng, src, pth, tgt, mask, target), keep=1.0)
            eng.adam_step()
    a, b = lazy.export_params(), dense.export_params()
    for k in O.PARAM_NAMES:
        assert np.abs(a[k] - b[k]).max() < 2e-6, k
    for name, cols in (("tok", (0, 2)), ("path", (1,))):
        counts = np.zeros(a[name].shape[0], dtype=np.int64)
        for batch in batches:
            for c in cols:
                np.add.at(counts, batch[c][batch[3] > 0], 1)
        once = counts <= 1                                  # never referenced, or by exactly one context: no atomic-order freedom
        assert once.sum() > 50 and np.array_equal(a[name][once], b[name][once]), name
    import torch
    assert torch.allclose(lazy.adam_m["tok"], dense.adam_m["tok"], atol=1e-7)
    assert torch.allclose(lazy.adam_v["path"], dense.adam_v["path"], atol=1e-9)


def test_sampled_softmax_updates_target_rows_lazily_and_exactly():
    """BASELINE config 3 with lazy Adam: only the B + S target rows a sampled-softmax step reads are touched
    (TF1 would apply the IndexedSlices gradient as a dense Adam step over all rows -- the deferred replay gives the
    same result).  Five sampled steps, then a full-softmax step (the table leaves the lazy set), against a dense
    engine doing the same."""
    import torch
    S = 7
    lazy, params0 = make_engine(DIMS, max_batch=B)
    dense, _ = make_engine(DIMS, max_batch=B, params=params0)
    lazy.set_option("lazy_adam", 1)
    lazy.set_option("adam_sweep_period", 4)
    rng = np.random.default_rng(9)
    for s in range(6):
        src, pth, tgt, mask, target = O.synthetic_batch(DIMS, B, seed=700 + s)
        sampled = O.log_uniform_sample(rng, S, DIMS.target_vocab)
        lq_t, lq_s = O.log_uniform_logq(target, S, DIMS.target_vocab), O.log_uniform_logq(sampled, S, DIMS.target_vocab)
        losses = []
        for eng in (lazy, dense):
            d = dev_batch(eng, src, pth, tgt, mask, target)
            if s < 5:
                l = eng.sampled_train_step(*d, eng.to_device(sampled, torch.int32), eng.to_device(lq_t, torch.float32),
                                           eng.to_device(lq_s, torch.float32))
            else:
                l = eng.train_step(*d, keep=1.0)
            losses.append(float(l.cpu()[0]))
            eng.adam_step()
        assert abs(losses[0] - losses[1]) < 1e-5, (s, losses)
    a, b = lazy.export_params(), dense.export_params()
    for k in O.PARAM_NAMES:
        assert np.abs(a[k] - b[k]).max() < 2e-6, k
    assert torch.allclose(lazy.adam_m["tgt"], dense.adam_m["tgt"], atol=1e-7)
    assert torch.allclose(lazy.adam_v["tgt"], dense.adam_v["tgt"], atol=1e-9)


@pytest.mark.parametrize("period", [0, 32])
@pytest.mark.parametrize("embed_dim", [32, 128, 256])
def test_long_idle_rows_replay_exactly(embed_dim, period):
    """Rows touched once and then left alone for 300 steps (the "theta rests" exit of replay_row fires after ~150 idle
    steps; with the sweep they are replayed in slices of <= 32 steps, without it in one go at the flush), and rows
    never touched at all, against the dense engine: bit for bit.  embed_dim 32 / 256 cover rows narrower and wider
    than one 128-column slice of the warp."""
    import torch
    dims = O.Dims(token_vocab=1501, path_vocab=701, target_vocab=101, embed_dim=embed_dim, code_dim=64, max_contexts=6)
    steps = 300
    first = O.synthetic_batch(dims, B, seed=900)
    later = O.synthetic_batch(dims, B, seed=901)
    later = tuple(np.where(a > 0, 1 + (a % 40), a).astype(a.dtype) if i < 3 else a for i, a in enumerate(later))   # rows 1..40 only
    lazy, params0 = make_engine(dims, max_batch=B)
    dense, _ = make_engine(dims, max_batch=B, params=params0)
    lazy.set_option("lazy_adam", 1)
    lazy.set_option("adam_sweep_period", period)
    for s in range(steps):
        src, pth, tgt, mask, target = first if s == 0 else later
        for eng in (lazy, dense):
            eng.train_step(*dev_batch(eng, src, pth, tgt, mask, target), keep=1.0)
            eng.adam_step()
    a, b = lazy.export_params(), dense.export_params()
    for name, cols in (("tok", (0, 2)), ("path", (1,))):
        counts = np.zeros(a[name].shape[0], dtype=np.int64)
        for c in cols:
            np.add.at(counts, first[c][first[3] > 0], 1)
        idle = counts == 1
        idle[:41] = False                                   # rows the later batches keep using
        assert idle.sum() > 20
        assert np.array_equal(a[name][idle].view(np.uint32), b[name][idle].view(np.uint32)), name
        assert np.abs(a[name][idle] - params0[name][idle]).max() > 1e-4          # they did move, then rested
        never = counts == 0
        never[:41] = False
        assert np.array_equal(a[name][never], params0[name][never])
        sel = torch.from_numpy(idle | never).to(lazy.dev)          # rows free of atomic-order effects: slots agree bit for bit too
        assert torch.equal(lazy.adam_m[name][sel], dense.adam_m[name][sel])
        assert torch.equal(lazy.adam_v[name][sel], dense.adam_v[name][sel])
    # the exit can be switched off; the result is the same
    assert lazy.get_option("adam_rest_shortcut") == 1
