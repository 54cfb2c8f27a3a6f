"""CPU-side check of the drop-in boundary: libc2v_b200.so builds, loads, and exports every
symbol include/c2v_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

from code2vec_b200 import build as B
from code2vec_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "c2v_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c2v_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = B.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "missing symbol %s" % name
    # the Python binding types exactly the declared set
    assert sorted(E._SIGNATURES) == declared


def test_abi_version_and_dims_validation_without_gpu():
    lib = E.load_library()
    assert lib.c2v_abi_version() == 1
    good = E.c2v_dims(1001, 501, 1001, 32, 96, 20, 64, 10)
    assert lib.c2v_workspace_bytes(ctypes.byref(good)) > 0
    bad = E.c2v_dims(1001, 501, 1001, 30, 96, 20, 64, 10)
    assert lib.c2v_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"embed_dim" in lib.c2v_last_error(None)
    # no device here (or a wrong one): create must fail loudly, never fall back
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        rc = lib.c2v_create(ctypes.byref(good), 0, ctypes.byref(h))
        assert rc < 0 and not h.value
        assert len(lib.c2v_last_error(None)) > 0


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        E.PathAttentionEngine(E.EngineDims(1001, 501, 1001, 32, 96, 20, 64))


def test_every_engine_option_is_documented_in_the_header():
    """Every key c2v_set_option / c2v_get_option accept (the strcmp chain in engine.cu) is described in include/c2v_b200.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "code2vec_b200", "csrc", "engine.cu")).read()
    hdr = open(os.path.join(root, "include", "c2v_b200.h")).read()
    keys = sorted(set(re.findall(r'!strcmp\(key, "([a-z_0-9]+)"\)', src)))
    assert len(keys) >= 20
    missing = [k for k in keys if '"%s"' % k not in hdr]
    assert not missing, missing
