"""Structural properties of the reference's graph (tensorflow_model.py:197-265) that the oracle must have
whatever the numbers are -- checked in fp64 on random small problems (hypothesis picks the shapes).
They follow from the source alone, so they pin the restatement without TensorFlow:
  * examples are independent up to the batch mean of the loss  (:226-230: per-example bag -> code vector);
  * a bag is a SET of contexts: permuting its slots permutes attention and changes nothing else (:254-263);
  * what sits in a masked slot is irrelevant, to outputs and to every gradient  (:257-260, log(mask) = -inf);
  * a context that appears twice receives twice the gradient of its table rows (IndexedSlices are summed, A.3);
  * the token table is shared by source and target terminals  (:238,240 read the same variable).
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import path_attention_oracle as O

F64 = np.float64
SETTINGS = dict(max_examples=12, deadline=None)


def _problem(seed, B, C, d, D, T=23, P=17, Y=19):
    dims = O.Dims(token_vocab=T, path_vocab=P, target_vocab=Y, embed_dim=d, code_dim=D, max_contexts=C)
    params = {k: v.astype(F64) for k, v in O.init_params(dims, seed=seed).items()}
    src, pth, tgt, mask, target = O.synthetic_batch(dims, B, seed=seed + 1)
    mask[:, 0] = 1.0                                    # no all-masked bag (that case has its own test)
    src[:, 0] = np.maximum(src[:, 0], 1)
    return dims, params, src, pth, tgt, mask, target


shapes = st.tuples(st.integers(0, 10 ** 6), st.integers(2, 5), st.integers(2, 7), st.sampled_from([4, 8]), st.sampled_from([4, 12]))


@settings(**SETTINGS)
@given(shapes)
def test_examples_are_independent_and_the_loss_is_their_mean(s):
    seed, B, C, d, D = s
    dims, params, src, pth, tgt, mask, target = _problem(seed, B, C, d, D)
    loss, grads, aux = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=F64)
    per_loss, acc = [], {k: np.zeros_like(v) for k, v in params.items()}
    for b in range(B):
        sl = slice(b, b + 1)
        lb, gb, auxb = O.train_loss_and_grads(params, src[sl], pth[sl], tgt[sl], mask[sl], target[sl], dtype=F64)
        assert np.allclose(auxb["v"], aux["v"][sl], atol=1e-13) and np.allclose(auxb["alpha"], aux["alpha"][sl], atol=1e-13)
        per_loss.append(lb)
        for k in acc:
            acc[k] += gb[k] / B
    assert abs(loss - np.mean(per_loss)) < 1e-12
    for k in acc:
        assert np.allclose(grads[k], acc[k], atol=1e-12), k


@settings(**SETTINGS)
@given(shapes)
def test_a_bag_is_a_set_of_contexts(s):
    seed, B, C, d, D = s
    dims, params, src, pth, tgt, mask, target = _problem(seed, B, C, d, D)
    perm = np.random.default_rng(seed).permutation(C)
    loss, grads, aux = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=F64)
    loss2, grads2, aux2 = O.train_loss_and_grads(params, src[:, perm], pth[:, perm], tgt[:, perm], mask[:, perm], target, dtype=F64)
    assert abs(loss - loss2) < 1e-12 and np.allclose(aux["v"], aux2["v"], atol=1e-13)
    assert np.allclose(aux["alpha"][:, perm], aux2["alpha"], atol=1e-14)
    for k in grads:
        assert np.allclose(grads[k], grads2[k], atol=1e-12), k


@settings(**SETTINGS)
@given(shapes)
def test_masked_slots_carry_no_information(s):
    seed, B, C, d, D = s
    dims, params, src, pth, tgt, mask, target = _problem(seed, B, C, d, D)
    mask[:, -1] = 0.0                                    # force at least one masked slot per bag
    rng = np.random.default_rng(seed + 7)
    src2, pth2, tgt2 = src.copy(), pth.copy(), tgt.copy()
    hidden = mask == 0
    src2[hidden] = rng.integers(0, dims.token_vocab, hidden.sum())
    pth2[hidden] = rng.integers(0, dims.path_vocab, hidden.sum())
    tgt2[hidden] = rng.integers(0, dims.token_vocab, hidden.sum())
    a = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=F64)
    b = O.train_loss_and_grads(params, src2, pth2, tgt2, mask, target, dtype=F64)
    assert a[0] == b[0] and np.array_equal(a[2]["v"], b[2]["v"]) and np.array_equal(a[2]["alpha"], b[2]["alpha"])
    assert np.all(a[2]["alpha"][hidden] == 0.0)
    for k in a[1]:
        assert np.array_equal(a[1][k], b[1][k]), k      # exact: masked contexts contribute exact zeros


def test_repeated_contexts_sum_their_gradients_and_the_token_table_is_shared():
    dims, params, src, pth, tgt, mask, target = _problem(3, 1, 4, 4, 8)
    # one bag, two distinct contexts, the first one present twice: [c0, c0, c1, pad]
    src[0], pth[0], tgt[0], mask[0] = [5, 5, 9, 0], [3, 3, 4, 0], [7, 7, 5, 0], [1, 1, 1, 0]
    loss, g, aux = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=F64)
    assert abs(aux["alpha"][0, 0] - aux["alpha"][0, 1]) < 1e-15          # identical contexts, identical attention
    # writing the duplicate once is a DIFFERENT bag (its softmax weights change), so duplicates are not collapsed
    src1, pth1, tgt1, mask1 = np.array([[5, 9, 0, 0]]), np.array([[3, 4, 0, 0]]), np.array([[7, 5, 0, 0]]), np.array([[1., 1., 0., 0.]])
    _, _, aux1 = O.train_loss_and_grads(params, src1, pth1, tgt1, mask1, target, dtype=F64)
    assert not np.allclose(aux["v"], aux1["v"])
    # its table rows receive the SUM over both occurrences: the analytic gradient of path row 3 (touched by slots 0
    # and 1 only) matches the finite difference of the loss in that row
    eps = 1e-6
    def loss_of(delta):
        p2 = {k: v.copy() for k, v in params.items()}
        p2["path"][3] += delta
        return O.train_loss_and_grads(p2, src, pth, tgt, mask, target, dtype=F64)[0]
    direction = np.random.default_rng(0).standard_normal(dims.embed_dim)
    fd = (loss_of(eps * direction) - loss_of(-eps * direction)) / (2 * eps)
    assert abs(fd - g["path"][3] @ direction) < 1e-8
    # token row 5 is read as a SOURCE terminal (slots 0, 1) and as a TARGET terminal (slot 2): one shared table
    def loss_tok(delta):
        p2 = {k: v.copy() for k, v in params.items()}
        p2["tok"][5] += delta
        return O.train_loss_and_grads(p2, src, pth, tgt, mask, target, dtype=F64)[0]
    fd = (loss_tok(eps * direction) - loss_tok(-eps * direction)) / (2 * eps)
    assert abs(fd - g["tok"][5] @ direction) < 1e-8
    assert np.abs(g["tok"][5]).max() > 0 and np.all(g["tok"][11] == 0)   # an unreferenced row gets exactly nothing


def test_attention_backward_needs_no_first_pass_over_the_bag():
    """attn_bwd_kernel (csrc/kernels.cuh) reads H once: the bag-wide term t = sum_c alpha_c (h_c . dv) of the softmax
    backward equals v . dv with the code vector v = sum_c alpha_c h_c the forward pass already produced.  Checked on the
    oracle's own forward quantities, ragged bags included, in fp32 -- the two summation orders differ by rounding only."""
    dims = O.Dims(token_vocab=301, path_vocab=211, target_vocab=97, embed_dim=12, code_dim=36, max_contexts=17)
    params = O.init_params(dims, seed=11)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, 29, seed=12)
    v, alpha, cache = O.forward(params, src, pth, tgt, mask)
    logits = O.logits_of(params, v)
    _, dv = O.backward(params, src, pth, tgt, mask, target, cache, v, logits)
    h = cache.h.reshape(src.shape[0], src.shape[1], -1)
    t_two_pass = (alpha * np.einsum("bcd,bd->bc", h, dv)).sum(axis=1)
    t_one_pass = np.einsum("bd,bd->b", v, dv)
    assert np.abs(t_one_pass - t_two_pass).max() <= 4e-7 * max(np.abs(t_two_pass).max(), 1e-30) + 1e-12
    # ... and the gradient of the attention vector built from it is the oracle's
    dalpha = np.einsum("bcd,bd->bc", h, dv)
    dz = alpha * (dalpha - t_one_pass[:, None])
    g_a = np.einsum("bc,bcd->d", dz, h)
    ref, _ = O.backward(params, src, pth, tgt, mask, target, cache, v, logits)
    assert np.abs(g_a - ref["a"]).max() <= 1e-5 * np.abs(ref["a"]).max() + 1e-12
