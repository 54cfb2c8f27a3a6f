"""code2vec_b200: B200-native (sm_100a) backend for code2vec's path-attention hot path.

Layout: ``csrc/`` CUDA kernels + the C ABI (include/c2v_b200.h) -> ``libc2v_b200.so``;
``engine.py`` ctypes binding + storage; the host-side mirror of the reference's model / config /
reader interface lives beside it (``config.py``, ``vocabularies.py``, ``path_context_reader.py``,
``model_base.py``, ``b200_model.py``).
"""
__version__ = "0.1.0"
