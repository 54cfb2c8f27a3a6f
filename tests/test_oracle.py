"""Pins for the CPU oracle: the reference has no golden vectors, so the oracle
is pinned by (i) torch autograd on an independent statement of the forward graph, (ii) float64
finite differences, (iii) semantic properties stated in the reference source."""
import numpy as np

from oracle import path_attention_oracle as O
from oracle import torch_crosscheck as TC

TINY = O.Dims(token_vocab=53, path_vocab=31, target_vocab=47, embed_dim=8, code_dim=24, max_contexts=7)


def _problem(dims=TINY, B=5, seed=7, dtype=np.float64):
    params = O.init_params(dims, seed=seed, dtype=dtype)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, B, seed=seed + 1)
    # force duplicates (same row hit by src and tgt, and by several contexts)
    src[0, 0] = tgt[0, 0] = src[1, 0] = 3
    pth[0, 0] = pth[1, 0] = 2
    return params, src, pth, tgt, mask, target


def test_backward_matches_torch_autograd_fp64():
    import torch
    params, src, pth, tgt, mask, target = _problem()
    loss, grads, aux = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=np.float64)
    tl, tg, taux = TC.loss_and_grads(params, src, pth, tgt, mask, target, dtype=torch.float64)
    assert abs(loss - tl) < 1e-12
    for k in O.PARAM_NAMES:
        np.testing.assert_allclose(grads[k], tg[k], rtol=0, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(aux["v"], taux["v"], atol=1e-14)
    np.testing.assert_allclose(aux["alpha"], taux["alpha"], atol=1e-14)


def test_backward_with_dropout_matches_torch_autograd_fp64():
    import torch
    params, src, pth, tgt, mask, target = _problem()
    B, C = src.shape
    dm = O.dropout_keep_mask(seed=99, step=3, n_rows=B * C, ctx_dim=TINY.ctx_dim, keep=0.75)
    assert 0.6 < dm.mean() < 0.9
    loss, grads, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target, keep=0.75,
                                            dropout_mask=dm, dtype=np.float64)
    tl, tg, _ = TC.loss_and_grads(params, src, pth, tgt, mask, target, keep=0.75, dropout_mask=dm,
                                  dtype=torch.float64)
    assert abs(loss - tl) < 1e-12
    for k in O.PARAM_NAMES:
        np.testing.assert_allclose(grads[k], tg[k], rtol=0, atol=1e-13, err_msg=k)


def test_backward_finite_differences_fp64():
    params, src, pth, tgt, mask, target = _problem(B=3)
    loss, grads, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=np.float64)
    rng = np.random.default_rng(0)
    eps = 1e-6
    for k in O.PARAM_NAMES:
        flat = params[k].reshape(-1)
        # probe touched entries (non-zero gradient) and a few random ones
        nz = np.flatnonzero(grads[k].reshape(-1))
        probe = np.concatenate([rng.choice(nz, size=min(6, nz.size), replace=False),
                                rng.integers(0, flat.size, size=3)])
        for i in probe:
            old = flat[i]
            flat[i] = old + eps
            lp, _, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=np.float64)
            flat[i] = old - eps
            lm, _, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target, dtype=np.float64)
            flat[i] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - grads[k].reshape(-1)[i]) < 1e-8, (k, i, fd, grads[k].reshape(-1)[i])


def test_masked_contexts_are_exact_zeros():
    """tensorflow_model.py:257-260: log(0) = -inf  =>  alpha == 0 exactly for padded slots, and the
    padded slots contribute nothing to any gradient (row 0 only collects exact zeros)."""
    params, src, pth, tgt, mask, target = _problem(dtype=np.float32)
    # make row 0 unused by valid contexts
    src[mask > 0] = np.maximum(src[mask > 0], 1)
    loss, grads, aux = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
    assert np.all(aux["alpha"][mask == 0] == 0.0)
    np.testing.assert_allclose(aux["alpha"].sum(axis=1), 1.0, atol=1e-6)
    assert np.all(grads["tok"][0] == 0.0) and np.all(grads["path"][0] == 0.0)


def test_all_masked_bag_is_nan_like_tf():
    """tf.nn.softmax over a row of -inf is NaN [TF-lib]; the reader filters such rows in
    train/eval (path_context_reader.py:153-177) but not in predict (:144-145)."""
    params, src, pth, tgt, mask, target = _problem(dtype=np.float32)
    mask[2, :] = 0
    v, alpha, _ = O.forward(params, src, pth, tgt, mask)
    assert np.all(np.isnan(alpha[2])) and np.all(np.isnan(v[2]))
    assert np.all(np.isfinite(v[[0, 1, 3, 4]]))


def test_loss_at_init_is_log_vocab():
    dims = O.Dims(1001, 501, 1001, 32, 96, 20)
    params = O.init_params(dims)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, 64)
    loss, _, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
    assert abs(loss - np.log(dims.target_vocab)) < 5e-2


def test_topk_ties_prefer_lower_index_and_sorted_desc():
    s = np.array([[0.5, 2.0, 2.0, -1.0, 2.0, 0.5]], dtype=np.float32)
    idx, vals = O.top_k(s, 4)
    assert idx.tolist() == [[1, 2, 4, 0]]
    assert vals.tolist() == [[2.0, 2.0, 2.0, 0.5]]
    idx, vals = O.top_k(s, 10)                      # k = min(10, Y)  tensorflow_model.py:299-300
    assert idx.shape == (1, 6)
    p = O.softmax_over_k(vals)
    np.testing.assert_allclose(p.sum(axis=1), 1.0, rtol=1e-6)


def test_adam_first_step_is_lr_sign_and_dense_decay():
    """TF1 Adam: after step 1 every touched entry moves by ~lr*sign(g); on step 2 with a zero
    gradient the entry still moves (non-lazy sparse semantics)."""
    params = {k: np.zeros(s, np.float32) for k, s in
              dict(tok=(4, 2), path=(3, 2), tgt=(5, 6), W=(6, 6), a=(6,)).items()}
    grads = {k: np.zeros_like(p) for k, p in params.items()}
    grads["tok"][1, 0] = 0.3
    grads["tok"][2, 1] = -4.0
    m = {k: np.zeros_like(p) for k, p in params.items()}
    v = {k: np.zeros_like(p) for k, p in params.items()}
    O.adam_step(params, grads, m, v, t=1)
    np.testing.assert_allclose(params["tok"][1, 0], -1e-3, rtol=1e-5)
    np.testing.assert_allclose(params["tok"][2, 1], +1e-3, rtol=1e-5)
    assert params["tok"][0, 0] == 0.0
    before = params["tok"][1, 0]
    zero = {k: np.zeros_like(p) for k, p in params.items()}
    O.adam_step(params, zero, m, v, t=2)
    assert params["tok"][1, 0] < before          # still moving with g == 0


def test_adam_matches_torch_cpu_trainer():
    dims = TINY
    params = O.init_params(dims, seed=5)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, 6, seed=11)
    tr = TC.TorchCpuTrainer({k: v.copy() for k, v in params.items()}, threads=1)
    m = {k: np.zeros_like(p) for k, p in params.items()}
    vv = {k: np.zeros_like(p) for k, p in params.items()}
    for t in (1, 2, 3):
        loss, grads, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
        O.adam_step(params, grads, m, vv, t)
        tl = tr.train_step(src, pth, tgt, mask, target)
        assert abs(loss - tl) < 2e-5
    tp = tr.numpy_params()
    for k in O.PARAM_NAMES:
        np.testing.assert_allclose(params[k], tp[k], atol=2e-5, err_msg=k)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors):
    counter = key = 0  and  counter = key = 0xffffffff."""
    r = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    r = O.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_sampled_softmax_grads_match_autograd():
    import torch
    dims = TINY
    params = O.init_params(dims, seed=3, dtype=np.float64)
    rng = np.random.default_rng(1)
    B, S = 6, 5
    v = rng.standard_normal((B, dims.code_dim))
    target = rng.integers(1, dims.target_vocab, size=B).astype(np.int32)
    sampled = O.log_uniform_sample(rng, S, dims.target_vocab)
    sampled[0] = target[2]                      # an accidental hit
    lq_t = O.log_uniform_logq(target, S, dims.target_vocab)
    lq_s = O.log_uniform_logq(sampled, S, dims.target_vocab)
    loss, dv, g_tgt, logits = O.sampled_softmax_loss_and_grads(params, v, target, sampled, lq_t, lq_s,
                                                               dtype=np.float64)
    tv = torch.tensor(v, requires_grad=True)
    ty = torch.tensor(params["tgt"], requires_grad=True)
    lt = (tv * ty[torch.as_tensor(target, dtype=torch.long)]).sum(1) - torch.tensor(lq_t, dtype=torch.float64)
    ls = tv @ ty[torch.as_tensor(sampled, dtype=torch.long)].t() - torch.tensor(lq_s, dtype=torch.float64)[None]
    hit = torch.tensor(sampled[None, :] == target[:, None])
    ls = torch.where(hit, torch.tensor(-1e9, dtype=torch.float64), ls)
    lg = torch.cat([lt[:, None], ls], 1)
    L = torch.nn.functional.cross_entropy(lg, torch.zeros(B, dtype=torch.long), reduction="sum") / B
    L.backward()
    assert abs(float(L.detach()) - loss) < 1e-12
    np.testing.assert_allclose(dv, tv.grad.numpy(), atol=1e-12)
    np.testing.assert_allclose(g_tgt, ty.grad.numpy(), atol=1e-12)


def test_keras_numerics_in_the_oracle():
    """Keras initialiser ranges and full-vocabulary softmax scores."""
    dims = O.Dims(token_vocab=500, path_vocab=300, target_vocab=700, embed_dim=16, code_dim=48, max_contexts=9)
    p = O.keras_init_params(dims, seed=1)
    assert np.abs(p["tok"]).max() <= 0.05 and np.abs(p["path"]).max() <= 0.05 and np.abs(p["a"]).max() <= 0.05
    assert np.abs(p["tgt"]).max() <= np.sqrt(6.0 / (48 + 700)) + 1e-7 and p["tgt"].shape == (700, 48)
    src, pth, tgt, mask, _ = O.synthetic_batch(dims, 6, seed=2)
    idx, probs, _, _, scores = O.evaluate_topk(p, src, pth, tgt, mask, k=5, normalize=2)
    e = np.exp(scores.astype(np.float64) - scores.max(1, keepdims=True))
    full = e / e.sum(1, keepdims=True)
    assert np.abs(probs - np.take_along_axis(full, idx, 1)).max() < 1e-7
    idx1, soft_k, _, _, _ = O.evaluate_topk(p, src, pth, tgt, mask, k=5, normalize=1)
    assert np.array_equal(idx, idx1) and np.allclose(soft_k.sum(1), 1.0, atol=1e-6) and np.all(probs.sum(1) < 1.0)
