"""The literal drop-in seam on the GPU: the faiss-API shim backed by the B200 engine under the REAL HF ``datasets``
chain the reference uses (``Dataset.add_faiss_index`` / ``Dataset.search_batch``, called from
meerqat/ir/search.py:245 and :146), against the same chain backed by the CPU oracle and against the
oracle-independent scikit-learn referee.  No reference checkout is needed: ``datasets`` is the dependency the
reference itself imports."""
import numpy as np
import pytest

from dropin import install_faiss_shim, mips_referee
from oracle import flat_ip

pytestmark = pytest.mark.gpu


def _search_through_datasets(faiss, backend, dtype, X, Q, factory, k):
    import datasets
    faiss.set_backend(backend, dtype)
    try:
        ds = datasets.Dataset.from_dict({"emb": X.tolist()})
        ds.add_faiss_index(column="emb", index_name="i", string_factory=factory, metric_type=faiss.METRIC_INNER_PRODUCT)
        res = ds.search_batch("i", Q, k)
        index = ds.get_index("i").faiss_index
        engine = getattr(index, "_engine", None)
        assert (engine is not None) == (backend == "b200"), "the shim did not route to the requested backend"
        return np.asarray(res.total_scores), np.asarray(res.total_indices), index.ntotal
    finally:
        faiss.set_backend("oracle", "f32")


@pytest.mark.parametrize("factory", ["Flat", "L2norm,Flat"])
@pytest.mark.parametrize("dtype", ["f32", "f32x3", "bf16"])
def test_engine_behind_datasets_faiss_chain(dtype, factory):
    faiss = install_faiss_shim()
    rng = np.random.default_rng(23)
    n, d, nq, k = 6000, 64, 37, 100
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if dtype == "bf16" and factory == "Flat":
        X, Q = flat_ip.bf16_round(X), flat_ip.bf16_round(Q)  # "same quantised inputs" (SURVEY §8c): rounding is then a no-op
    Dg, Ig, ntotal = _search_through_datasets(faiss, "b200", dtype, X, Q, factory, k)
    Do, Io, _ = _search_through_datasets(faiss, "oracle", "f32", X, Q, factory, k)
    assert ntotal == n and Dg.shape == (nq, k) and Ig.shape == (nq, k) and Dg.dtype == np.float32
    assert Ig.dtype.kind == "i" and np.all(np.diff(Dg, axis=1) <= 0)
    Xr, Qr = (flat_ip.l2norm(X), flat_ip.l2norm(Q)) if factory == "L2norm,Flat" else (X, Q)
    Dr, Ir = mips_referee(Qr, Xr, k)
    overlap_o = np.mean([len(set(a) & set(b)) / k for a, b in zip(Ig, Io)])
    overlap_r = np.mean([len(set(a) & set(b)) / k for a, b in zip(Ig, Ir)])
    if dtype in ("f32", "f32x3"):
        # fp32 bar of north_star: ids equal except between (near-)tied scores, scores within 1e-5 relative
        assert np.allclose(Dg, Do, rtol=1e-5, atol=1e-6) and np.allclose(Dg, Dr, rtol=1e-5, atol=1e-6)
        assert overlap_o >= 0.999 and overlap_r >= 0.999
        for I_ref in (Io, Ir):
            differ = Ig != I_ref
            assert not differ.any() or np.abs(Dg - Dr)[differ].max() <= 1e-5 * np.abs(Dr).max()
    elif factory == "Flat":
        # bf16 on the same quantised inputs: scores within 2e-3 relative, top-100 overlap >= 99.9 %
        assert np.allclose(Dg, Do, rtol=2e-3, atol=1e-6) and np.allclose(Dg, Dr, rtol=2e-3, atol=1e-6)
        assert overlap_o >= 0.999 and overlap_r >= 0.999
    else:
        # bf16 + L2norm: the engine rounds AFTER normalising, the fp32 chain never rounds, so this difference is the
        # storage type's quantisation loss (|error| <= 2^-9 per unit-vector element pair), not kernel error
        assert np.abs(Dg - Do).max() <= 4e-3 and overlap_o >= 0.97 and overlap_r >= 0.97


def test_shim_chain_rejects_what_datasets_rejects():
    """Misuse surfaces through the same datasets checks the reference relies on (datasets/search.py:380-381, 713-733)."""
    import datasets
    faiss = install_faiss_shim()
    faiss.set_backend("b200", "bf16")
    try:
        ds = datasets.Dataset.from_dict({"emb": np.eye(8, dtype=np.float32).tolist()})
        ds.add_faiss_index(column="emb", index_name="i", string_factory="Flat", metric_type=0)
        with pytest.raises(ValueError):
            ds.search_batch("i", np.zeros(8, np.float32), 3)       # 1-D queries
        with pytest.raises(Exception):
            ds.search_batch("missing", np.zeros((1, 8), np.float32), 3)
        s, i = ds.search_batch("i", np.eye(8, dtype=np.float32)[:2], 3)
        assert [int(r[0]) for r in i] == [0, 1]
    finally:
        faiss.set_backend("oracle", "f32")
