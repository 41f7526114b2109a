"""Test-only stand-in for the 6-symbol slice of the FAISS Python API that HF ``datasets`` touches
(datasets/search.py:276-313, 369-385): METRIC_*, IndexFlat, index_factory, .add, .search, .ntotal.

It lets the reference's UNMODIFIED meerqat/ir/search.py drive either the CPU oracle (default) or
the B200 engine (``set_backend("b200")``) through the real datasets -> faiss call chain — the
literal drop-in seam.  TEST INFRASTRUCTURE: never imported by viquae_b200.
"""
import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
_BACKEND = {"name": "oracle", "dtype": "f32"}


def set_backend(name, dtype="f32"):
    _BACKEND["name"] = name
    _BACKEND["dtype"] = dtype


class _FlatIP:
    def __init__(self, d, l2norm=False):
        self.d = d
        self.l2norm = l2norm
        self.verbose = False
        self.is_trained = True
        self._rows = []
        self._engine = None

    @property
    def ntotal(self):
        if self._engine is not None:
            return self._engine.ntotal
        return sum(len(r) for r in self._rows)

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        if _BACKEND["name"] == "b200":
            if self._engine is None:
                from viquae_b200.index import B200FlatIndex
                self._engine = B200FlatIndex(d=self.d, dtype=_BACKEND["dtype"], device=0,
                                             string_factory="L2norm,Flat" if self.l2norm else "Flat")
            self._engine.add(x)
        else:
            from oracle import flat_ip
            self._rows.append(flat_ip.l2norm(x) if self.l2norm else x)

    def search(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if self._engine is not None:
            res = self._engine.search_batch(q, k)
            return np.asarray(res.total_scores, np.float32), np.asarray(res.total_indices, np.int64)
        from oracle import flat_ip
        X = np.concatenate(self._rows) if self._rows else np.zeros((0, self.d), np.float32)
        return flat_ip.flat_ip_search(q, X, k, l2norm_queries=self.l2norm)


def IndexFlat(d, metric=METRIC_L2):
    if metric != METRIC_INNER_PRODUCT:
        raise NotImplementedError("shim: inner product only")
    return _FlatIP(d)


def IndexFlatIP(d):
    return _FlatIP(d)


def index_factory(d, string_factory, metric=METRIC_L2):
    if metric != METRIC_INNER_PRODUCT:
        raise NotImplementedError("shim: inner product only")
    parts = [p.strip() for p in string_factory.split(",")]
    if parts == ["Flat"]:
        return _FlatIP(d)
    if parts == ["L2norm", "Flat"]:
        return _FlatIP(d, l2norm=True)
    raise NotImplementedError(f"shim: unsupported factory {string_factory!r}")
