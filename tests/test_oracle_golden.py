"""The committed fixture tests/golden/path_attention_golden.npz against (a) the CPU oracle, so a
change of the oracle's arithmetic cannot go unnoticed (CPU), and (b) the CUDA path through the C
ABI on the same inputs (GPU)."""
import os

import numpy as np
import pytest

from oracle import path_attention_oracle as O
from tests.golden import make_golden as MG

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "path_attention_golden.npz"))
DIMS = O.Dims(**MG.DIMS)


def _params():
    return {k: G["param_" + k] for k in O.PARAM_NAMES}


def test_oracle_reproduces_committed_fixture():
    fresh = MG.build()
    for key in G.files:
        a, b = G[key], fresh[key]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), key
        else:
            np.testing.assert_allclose(a, b, rtol=2e-6, atol=2e-7, err_msg=key)


def test_fixture_is_self_consistent():
    assert np.all(G["attention"][G["mask"] == 0] == 0)
    np.testing.assert_allclose(G["attention"].sum(axis=1), 1.0, atol=1e-6)
    assert abs(float(G["loss"]) - np.log(DIMS.target_vocab)) < 0.1
    assert np.all(np.diff(G["topk_val"], axis=1) <= 0)
    assert np.all(G["grad_tok"][0] == 0) and np.all(G["grad_path"][0] == 0)


@pytest.mark.gpu
def test_cuda_path_matches_committed_fixture():
    from tests.util import dev_batch, make_engine, rel_err
    eng, _ = make_engine(DIMS, max_batch=MG.B, params=_params())
    d = dev_batch(eng, G["src"], G["pth"], G["tgt"], G["mask"], G["target"])
    code, attn = eng.forward(*d[:4])
    assert rel_err(code.cpu().numpy(), G["code_vectors"]) < 2e-5
    assert np.abs(attn.cpu().numpy() - G["attention"]).max() < 2e-6
    idx, val = eng.topk(code)
    gaps = np.abs(np.diff(G["topk_val"], axis=1)).min(axis=1)
    ok = gaps > 1e-6
    assert np.array_equal(idx.cpu().numpy()[ok], G["topk_idx"][ok])
    np.testing.assert_allclose(val.cpu().numpy(), G["topk_val"], atol=1e-5)
    loss = float(eng.train_step(*d).cpu()[0])
    assert abs(loss - float(G["loss"])) < 1e-4
    g = eng.export_grads()
    for k in O.PARAM_NAMES:
        assert rel_err(g[k], G["grad_" + k]) < 5e-5, k
    loss = float(eng.train_step(*d, keep=0.75, seed=2024, step=5).cpu()[0])
    assert abs(loss - float(G["loss_dropout"])) < 1e-4
    g = eng.export_grads()
    for k in O.PARAM_NAMES:
        assert rel_err(g[k], G["grad_dropout_" + k]) < 5e-5, k
    eng.load_params(_params())
    for _ in range(2):
        eng.train_step(*d)
        eng.adam_step()
    got = eng.export_params()
    for k in O.PARAM_NAMES:
        assert np.abs(got[k] - G["adam2_" + k]).max() < 3e-5, k
