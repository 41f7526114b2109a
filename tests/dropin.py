"""Helpers of the literal drop-in tests: the faiss-API shim (tests/shims/faiss) under the REAL HF ``datasets``
call chain (``Dataset.add_faiss_index`` -> ``FaissIndex.add_vectors`` -> ``faiss.index_factory`` / ``.add``;
``Dataset.search_batch`` -> ``FaissIndex.search_batch`` -> ``faiss_index.search``, datasets/search.py:255-313, 369-385),
and an oracle-independent referee for exact inner-product top-k.  TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def install_faiss_shim():
    """Makes ``import faiss`` resolve to the shim and tells ``datasets`` that faiss exists (it probes
    ``find_spec("faiss")`` once at import time, datasets/search.py:30)."""
    shim_dir = os.path.join(HERE, "shims")
    if shim_dir not in sys.path:
        sys.path.insert(0, shim_dir)
    import faiss
    import datasets.search as ds_search
    ds_search._has_faiss = True
    return faiss


def mips_referee(Q, X, k):
    """Exact maximum-inner-product top-k WITHOUT any code of ``oracle/``: scikit-learn's brute-force Euclidean
    nearest neighbours (float64) over the standard MIPS -> L2 augmentation x' = [x, sqrt(M^2 - |x|^2)], q' = [q, 0]
    (|q' - x'|^2 = |q|^2 + M^2 - 2 q.x, so the nearest neighbours are the largest inner products).  Scores are
    recomputed as float64 dot products of the returned rows; ties are ordered by lower id like the engine."""
    from sklearn.neighbors import NearestNeighbors

    Q64, X64 = np.asarray(Q, np.float64), np.asarray(X, np.float64)
    n2 = (X64 * X64).sum(1)
    M2 = float(n2.max()) * (1 + 1e-9) + 1e-30
    Xa = np.concatenate([X64, np.sqrt(np.maximum(M2 - n2, 0.0))[:, None]], axis=1)
    Qa = np.concatenate([Q64, np.zeros((len(Q64), 1))], axis=1)
    kk = min(len(X64), k + 8)  # a few extra neighbours so that ties at rank k can be ordered by id
    nn = NearestNeighbors(n_neighbors=kk, algorithm="brute", metric="euclidean").fit(Xa)
    _, idx = nn.kneighbors(Qa)
    D = np.einsum("qd,qkd->qk", Q64, X64[idx])
    order = np.lexsort((idx, -D), axis=1)
    idx = np.take_along_axis(idx, order, 1)[:, :k]
    D = np.take_along_axis(D, order, 1)[:, :k]
    return D, idx.astype(np.int64)
