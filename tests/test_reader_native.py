"""The native (C++) tensoriser against the pure-Python reader: identical rows, masks, filtering,
target strings and error behaviour on generated `.c2v` files, including padding, OOV words, empty
pieces, short contexts and malformed lines."""
import pickle

import numpy as np
import pytest

from code2vec_b200 import vocabularies as V
from code2vec_b200.config import Config
from code2vec_b200.path_context_reader import EstimatorAction, PathContextReader, load_native_tensoriser


class _Former:
    def to_model_input_form(self, t):
        return t

    def from_model_input_form(self, row):
        return row


def _setup(tmp_path, lines, separate=False, C=6, batch=4, epochs=1):
    rng = np.random.default_rng(0)
    prefix = str(tmp_path / "ds")
    tok = {"t%d" % i: int(rng.integers(1, 50)) for i in range(30)}
    pth = {str(100 + i): int(rng.integers(1, 50)) for i in range(20)}
    tgt = {"name|%d" % i: int(rng.integers(1, 50)) for i in range(10)}
    with open(prefix + ".dict.c2v", "wb") as f:
        for d in (tok, pth, tgt):
            pickle.dump(d, f)
        pickle.dump(len(lines), f)
    with open(prefix + ".train.c2v", "w") as f:
        f.write("".join(l + "\n" for l in lines))
    cfg = Config(set_defaults=True)
    cfg.VERBOSE_MODE = 0
    cfg.TRAIN_DATA_PATH_PREFIX = prefix
    cfg.TEST_DATA_PATH = prefix + ".train.c2v"
    cfg.MAX_CONTEXTS = C
    cfg.TRAIN_BATCH_SIZE = cfg.TEST_BATCH_SIZE = batch
    cfg.NUM_TRAIN_EPOCHS = epochs
    cfg.SHUFFLE_BUFFER_SIZE = 8
    cfg.SEPARATE_OOV_AND_PAD = separate
    cfg.READER_NUM_PARALLEL_BATCHES = 3
    return cfg, V.Code2VecVocabs(cfg)


def _random_lines(n, C, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        target = ["name|%d" % rng.integers(0, 12), "", "unknown"][int(rng.choice(3, p=[0.8, 0.05, 0.15]))]
        k = int(rng.integers(0, C + 1))
        ctxs = []
        for _ in range(k):
            kind = rng.integers(0, 10)
            s, p, t = "t%d" % rng.integers(0, 36), str(100 + rng.integers(0, 25)), "t%d" % rng.integers(0, 36)
            if kind == 0:
                ctxs.append("%s,%s" % (s, p))
            elif kind == 1:
                ctxs.append(",,")
            elif kind == 2:
                ctxs.append("%s,,%s" % (s, t))
            elif kind == 3:
                ctxs.append("zzz,999,yyy")
            else:
                ctxs.append("%s,%s,%s" % (s, p, t))
        out.append(" ".join([target] + ctxs + [""] * (C - k)))
    return out


pytestmark = pytest.mark.skipif(load_native_tensoriser() is None, reason="g++ build of the native tensoriser failed")


@pytest.mark.parametrize("separate", [False, True])
def test_native_evaluate_matches_python_reader(tmp_path, separate):
    C = 6
    lines = _random_lines(57, C, seed=3)
    cfg, vs = _setup(tmp_path, lines, separate=separate, C=C, batch=5)
    py = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=False)
    nat = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=True)
    a, b = list(py.get_dataset()), list(nat.get_dataset())
    assert len(a) == len(b) and len(a) > 3
    for x, y in zip(a, b):
        for name in ("path_source_token_indices", "path_indices", "path_target_token_indices", "context_valid_mask", "target_index"):
            u, v = getattr(x, name), getattr(y, name)
            assert u.dtype == v.dtype and np.array_equal(u, v), name
        assert list(x.target_string) == list(y.target_string)


def test_native_train_filter_and_epochs(tmp_path):
    C = 6
    lines = _random_lines(40, C, seed=5)
    cfg, vs = _setup(tmp_path, lines, C=C, batch=4, epochs=3)
    py = PathContextReader(vs, cfg, _Former(), EstimatorAction.Train, use_native=False, shuffle_seed=1)
    nat = PathContextReader(vs, cfg, _Former(), EstimatorAction.Train, use_native=True, shuffle_seed=1)

    def rows(ds):
        out = []
        for b in ds:
            for i in range(b.path_indices.shape[0]):
                out.append((int(b.target_index[i]),) + tuple(b.path_source_token_indices[i]) + tuple(b.path_indices[i])
                           + tuple(b.path_target_token_indices[i]) + tuple(b.context_valid_mask[i]))
        return out
    ra, rb = rows(py.get_dataset()), rows(nat.get_dataset())
    assert len(ra) == len(rb) and len(ra) % 3 == 0 and len(ra) > 0
    assert sorted(ra) == sorted(rb)                       # same multiset of rows (order is shuffled)
    assert all(r[0] > 0 for r in rb)                      # train drops OOV targets


def test_native_errors_match_python(tmp_path):
    C = 4
    good = " ".join(["name|1", "t1,100,t2", "", "", ""])
    for bad in ("name|1 t1,100,t2", " ".join(["name|1", "a,b,c,d", "", "", ""])):
        cfg, vs = _setup(tmp_path, [good, bad], C=C, batch=2)
        for native in (False, True):
            r = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=native)
            with pytest.raises(ValueError):
                list(r.get_dataset())


@pytest.mark.parametrize("threads", [1, 5])
def test_native_matches_python_on_a_larger_fuzz(tmp_path, threads):
    """2000 generated lines (every field shape of _random_lines), 20 contexts, with the line range split over
    1 and 5 parser threads: the pipelined lookups must give the Python reader's rows."""
    C = 20
    lines = _random_lines(2000, C, seed=11)
    cfg, vs = _setup(tmp_path, lines, C=C, batch=64)
    cfg.READER_NUM_PARALLEL_BATCHES = threads
    py = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=False)
    nat = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=True)
    a, b = list(py.get_dataset()), list(nat.get_dataset())
    assert len(a) == len(b) > 20
    for x, y in zip(a, b):
        for name in ("path_source_token_indices", "path_indices", "path_target_token_indices", "context_valid_mask", "target_index"):
            assert np.array_equal(getattr(x, name), getattr(y, name)), name
        assert list(x.target_string) == list(y.target_string)


def test_blank_lines_are_skipped_by_both_parsers(tmp_path):
    """A blank line is not a record: the Python statement skips "" / "\\n" and the native tensoriser does the same
    (whether a dataset loads must not depend on which of the two is in use)."""
    C = 6
    lines = _random_lines(23, C, seed=11)
    with_blanks = []
    for i, l in enumerate(lines):
        with_blanks.append(l)
        if i % 4 == 1:
            with_blanks.append("")
        if i % 9 == 2:
            with_blanks += ["", ""]
    cfg, vs = _setup(tmp_path, [""] + with_blanks + ["", ""], C=C, batch=5)
    (tmp_path / "plain").mkdir()
    cfg_plain, vs_plain = _setup(tmp_path / "plain", lines, C=C, batch=5)
    got = {}
    for native in (False, True):
        r = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=native)
        got[native] = list(r.get_dataset())
    plain = list(PathContextReader(vs_plain, cfg_plain, _Former(), EstimatorAction.Evaluate, use_native=True).get_dataset())
    assert len(got[False]) == len(got[True]) == len(plain)
    for x, y, z in zip(got[False], got[True], plain):
        for name in ("path_source_token_indices", "path_indices", "path_target_token_indices", "context_valid_mask", "target_index"):
            assert np.array_equal(getattr(x, name), getattr(y, name)) and np.array_equal(getattr(y, name), getattr(z, name)), name
        assert list(x.target_string) == list(y.target_string) == list(z.target_string)


def test_many_short_lines_raise_a_clear_error(tmp_path):
    """More (malformed, short) lines than the chunk's size allows well-formed records: a ValueError about the field
    count, not a nonsensical line number."""
    C = 8
    good = " ".join(["name|1", "t1,100,t2"] + [""] * (C - 1))
    cfg, vs = _setup(tmp_path, [good] + ["x"] * 50, C=C, batch=2)
    for native in (False, True):
        r = PathContextReader(vs, cfg, _Former(), EstimatorAction.Evaluate, use_native=native)
        with pytest.raises(ValueError, match="fields"):
            list(r.get_dataset())


@pytest.mark.parametrize("b,threads", [(1, 1), (64, 3), (200, 8), (300, 2)])
def test_native_pool_draw_is_the_numpy_draw(b, threads):
    """_RowPool.take through c2v_pool_take (one native call, row copies spread over threads) against its numpy statement:
    with the same random stream both return the same rows in the same order and leave the same pool behind -- draw after
    draw, into fresh arrays and into caller-provided (ring slot) buffers, down to an empty pool."""
    from code2vec_b200.path_context_reader import _RowPool
    lib = load_native_tensoriser()
    assert hasattr(lib, "c2v_pool_take")
    C, n = 7, 300
    rng = np.random.default_rng(5)
    rows = (rng.integers(0, 1000, size=(n, C), dtype=np.int32), rng.integers(0, 1000, size=(n, C), dtype=np.int32),
            rng.integers(0, 1000, size=(n, C), dtype=np.int32), rng.integers(0, 2, size=(n, C)).astype(np.float32),
            rng.integers(0, 1000, size=n, dtype=np.int32))
    pools = []
    for native in (None, lib):
        p = _RowPool(native=native, threads=threads)
        p.append(rows)
        assert (p.native is not None) == (native is not None)
        pools.append(p)
    r_np, r_nat = np.random.default_rng(9), np.random.default_rng(9)
    slot = tuple(np.empty((b,) + a.shape[1:], dtype=a.dtype) for a in rows)
    step = 0
    while pools[0].n > 0:
        k = min(b, pools[0].n)
        into = slot if step % 2 else None
        got_np = [x.copy() for x in pools[0].take(k, r_np)]
        got_nat = [x.copy() for x in pools[1].take(k, r_nat, out=into)]
        for x, y in zip(got_np, got_nat):
            assert x.dtype == y.dtype and np.array_equal(x, y)
        assert pools[0].n == pools[1].n
        for x, y in zip(pools[0].arrays, pools[1].arrays):
            assert np.array_equal(x[:pools[0].n], y[:pools[1].n])
        step += 1
    assert step == -(-n // b)


def test_native_pool_draw_rejects_a_bad_pick():
    lib = load_native_tensoriser()
    a = np.zeros((4, 3), dtype=np.int32)
    m = np.zeros((4, 3), dtype=np.float32)
    t = np.zeros(4, dtype=np.int32)
    o, om, ot = np.zeros((2, 3), dtype=np.int32), np.zeros((2, 3), dtype=np.float32), np.zeros(2, dtype=np.int32)
    for pick in ([1, 1], [0, 4], [-1, 2]):
        pk = np.asarray(pick, dtype=np.int64)
        assert lib.c2v_pool_take(a.ctypes.data, a.ctypes.data, a.ctypes.data, m.ctypes.data, t.ctypes.data, 4, 3, pk.ctypes.data, 2,
                                 o.ctypes.data, o.ctypes.data, o.ctypes.data, om.ctypes.data, ot.ctypes.data, 2) == -1
