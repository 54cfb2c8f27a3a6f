"""EXPERIMENTAL, not part of the default GPU suite: the long-idle shortcut of the lazy Adam row pass (engine option
"adam_rows_shortcut", off by default).  It has not run on a GPU yet -- its first version hung on
embed_dim < 128 -- so these tests only run with C2V_EXPERIMENTAL=1, under the suite's per-test timeout:
    C2V_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q
The scheme itself is proven bit-exact on a numpy model (tests/test_lazy_adam_model.py)."""
import os

import numpy as np
import pytest

from oracle import path_attention_oracle as O
from tests.util import dev_batch, make_engine

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("C2V_EXPERIMENTAL") != "1", reason="experimental kernels: set C2V_EXPERIMENTAL=1")]

DIMS = O.Dims(token_vocab=4001, path_vocab=2003, target_vocab=301, embed_dim=32, code_dim=96, max_contexts=10)
B = 8


@pytest.mark.parametrize("occ", [4, 5])
def test_long_idle_rows_take_the_underflow_fast_paths_and_stay_bit_exact(occ):
    """A row left alone for ~1000 steps: its m decays to exactly 0 (after which theta rests and only v decays) and
    then v does too.  adam_rows_kernel leaves the division / square-root loop at those points; the dense kernel
    grinds through every step.  Same bits."""
    import torch
    hp = dict(lr=1e-2, beta1=0.5, beta2=0.9, eps=1e-8)          # small betas: m is 0 after ~150 idle steps, v after ~900
    a = O.synthetic_batch(DIMS, B, seed=500)
    b = O.synthetic_batch(DIMS, B, seed=501)
    lazy, params0 = make_engine(DIMS, max_batch=B)
    dense, _ = make_engine(DIMS, max_batch=B, params=params0)
    lazy.set_option("lazy_adam", 1)
    lazy.set_option("adam_rows_shortcut", 1)
    lazy.set_option("adam_rows_occupancy", occ)
    n_idle = 1000
    plan = [a] + [b] * n_idle + [a]
    dev = {id(x): (dev_batch(lazy, *x), dev_batch(dense, *x)) for x in (a, b)}
    for batch in plan:
        dl, dd = dev[id(batch)]
        lazy.train_step(*dl, keep=1.0)
        lazy.adam_step(**hp)
        dense.train_step(*dd, keep=1.0)
        dense.adam_step(**hp)
    got, want = lazy.export_params(), dense.export_params()
    for name, cols in (("tok", (0, 2)), ("path", (1,))):
        in_a = np.zeros(got[name].shape[0], dtype=np.int64)
        in_b = np.zeros_like(in_a)
        for c in cols:
            np.add.at(in_a, a[c][a[3] > 0], 1)
            np.add.at(in_b, b[c][b[3] > 0], 1)
        idle = (in_a == 1) & (in_b == 0)                        # one context in batch a, none in b: no atomic-order freedom
        assert idle.sum() > 20, name
        assert np.array_equal(got[name][idle], want[name][idle]), name
        # the slots went all the way down: m is exactly 0 again, theta moved
        m_l, m_d = getattr(lazy, "adam_m")[name].cpu().numpy(), getattr(dense, "adam_m")[name].cpu().numpy()
        v_l, v_d = getattr(lazy, "adam_v")[name].cpu().numpy(), getattr(dense, "adam_v")[name].cpu().numpy()
        assert np.array_equal(m_l[idle], m_d[idle]) and np.array_equal(v_l[idle], v_d[idle]), name
        assert np.abs(got[name][idle] - params0[name][idle]).max() > 1e-3
    for k in O.PARAM_NAMES:
        assert np.abs(got[k] - want[k]).max() < 5e-6, k
