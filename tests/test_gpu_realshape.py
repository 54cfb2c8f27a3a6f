"""Parity at the REAL vocabulary sizes of BASELINE.json: the java14m shape (T = 1,301,137, P = 911,418,
Y = 261,246, d = 128, C = 200; configs[1]) and the large model shape (d = 256, D = 768; configs[4]) with the
java14m target vocabulary, in all three arithmetic modes, through the C ABI against the numpy oracle on
identical seeded inputs.  These sizes exercise what the small parity shapes cannot: the padded logits pitch
(261,246 is not a multiple of 64), 2042 log-sum-exp partial slots per row, split-K 18 / 48, the n-fastest tile
raster, and lazy-Adam bookkeeping over 2.2 M table rows.  The oracle costs a few seconds at B = 64.

Tolerances: loss within 1e-4 (BASELINE.json) in every mode; gradients 5e-5 relative (fp32 FFMA and 3xTF32) or
1e-2 (tf32); top-10 indices identical on every row whose top-11 logit gaps exceed 1e-6 (fp32-class modes; at
least 95 % of the rows must qualify) or 2e-3 (tf32)."""
import numpy as np
import pytest

from oracle import path_attention_oracle as O
from tests.util import dev_batch, make_engine, rel_err
from tests.test_gpu_parity import check_topk_rows

pytestmark = pytest.mark.gpu

JAVA14M = O.Dims(token_vocab=1301137, path_vocab=911418, target_vocab=261246, embed_dim=128, code_dim=384, max_contexts=200)
# configs[4]'s model shape with the java14m target vocabulary; the embedding tables are kept small so the oracle's
# dense gradient arrays stay cheap (their parity at 1.3 M rows is covered by JAVA14M)
LARGE_Y = O.Dims(token_vocab=30011, path_vocab=20011, target_vocab=261246, embed_dim=256, code_dim=768, max_contexts=200)

_cache = {}


def oracle_case(name):
    """(dims, B, params, batch, reference results) -- computed once per shape and shared by the three modes."""
    if name in _cache:
        return _cache[name]
    dims, B = (JAVA14M, 64) if name == "java14m" else (LARGE_Y, 32)
    params = O.init_params(dims, seed=4321)
    src, pth, tgt, mask, target = O.synthetic_batch(dims, B, seed=1234)
    src[0, 0] = tgt[0, 0] = src[1, 0] = 3                    # duplicates across src / tgt and across examples
    loss_ref, g_ref, _ = O.train_loss_and_grads(params, src, pth, tgt, mask, target)
    idx_ref, val_ref, v_ref, alpha_ref, scores = O.evaluate_topk(params, src, pth, tgt, mask, k=10)
    _cache[name] = (dims, B, params, (src, pth, tgt, mask, target), (loss_ref, g_ref, idx_ref, val_ref, v_ref, alpha_ref, scores))
    return _cache[name]


@pytest.mark.parametrize("math", [0, 1, 2])
@pytest.mark.parametrize("name", ["java14m", "large_y"])
def test_real_vocabulary_shape(name, math):
    dims, B, params, (src, pth, tgt, mask, target), (loss_ref, g_ref, idx_ref, val_ref, v_ref, alpha_ref, scores) = oracle_case(name)
    fp32_class = math != 1
    eng, _ = make_engine(dims, max_batch=B, params=params)
    eng.set_option("math_mode", math)
    d = dev_batch(eng, src, pth, tgt, mask, target)
    # ---- evaluate: code vectors, attention, top-10 ------------------------------------------------
    code, attn = eng.forward(*d[:4])
    assert rel_err(code.cpu().numpy(), v_ref) < (2e-5 if fp32_class else 3e-3)
    assert np.abs(attn.cpu().numpy() - alpha_ref).max() < (2e-6 if fp32_class else 2e-3)
    idx, val = eng.topk(code)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    if fp32_class:
        check_topk_rows(idx, idx_ref, scores, 10)
        assert np.abs(val - val_ref).max() < 1e-5
    else:
        check_topk_rows(idx, idx_ref, scores, 10, min_frac=0.0, gap_eps=2e-3)
        assert np.abs(val - val_ref).max() < 3e-3 * max(1.0, np.abs(val_ref).max())
    # ---- one train step: loss, all five gradients ----------------------------------------------------
    loss = float(eng.train_step(*d, keep=1.0).cpu()[0])
    assert abs(loss - loss_ref) < 1e-4, (loss, loss_ref)
    g = eng.export_grads()
    tol = 5e-5 if fp32_class else 1e-2
    for k in O.PARAM_NAMES:
        assert rel_err(g[k], g_ref[k]) < tol, (k, rel_err(g[k], g_ref[k]))
    touched = np.zeros(dims.token_vocab, bool)
    touched[src[mask > 0]] = True
    touched[tgt[mask > 0]] = True
    assert np.all(g["tok"][~touched] == 0.0)
    del g
    eng.close()


@pytest.mark.parametrize("math", [1, 2])
def test_real_shape_lazy_fused_steps_match_dense(math):
    """Three optimizer steps at the java14m shape: Trainer("single") (lazy embedding Adam with the sweep, target
    Adam in the dY epilogue) against the same engine run densely.  The target table must be bit-identical, the
    other tensors agree to atomic-order noise."""
    import torch
    from code2vec_b200.trainer import Trainer
    dims, B, params, _, _ = oracle_case("java14m")
    fast, _ = make_engine(dims, max_batch=B, params=params)
    dense, _ = make_engine(dims, max_batch=B, params=params)
    for eng in (fast, dense):
        eng.set_option("math_mode", math)
    tr = Trainer(fast, keep_prob=0.75, seed=3)
    for s in range(3):
        src, pth, tgt, mask, target = O.synthetic_batch(dims, B, seed=4000 + s)
        la = float(tr.step_device(*dev_batch(fast, src, pth, tgt, mask, target)).cpu()[0])
        lb = float(dense.train_step(*dev_batch(dense, src, pth, tgt, mask, target), keep=0.75, seed=3, step=s + 1).cpu()[0])
        dense.adam_step()
        assert abs(la - lb) < 1e-5, (s, la, lb)
    fast.sync_tables()
    torch.cuda.synchronize()
    assert torch.equal(fast.params["tgt"], dense.params["tgt"])
    for k in ("tok", "path", "W", "a"):
        assert (fast.params[k] - dense.params[k]).abs().max().item() < 2e-6, k
    fast.close(); dense.close()
