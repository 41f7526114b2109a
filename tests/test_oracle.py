"""The oracle against the golden fixtures produced by the reference's own Python layer, against
itself (numpy vs C restatement) and on the edge cases FAISS IndexFlatIP defines."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, flat_ip, fuse_oracle


def test_search_golden_matches_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_golden.npz"))
    meta = json.load(open(os.path.join(golden_dir, "search_golden.json")))
    k = meta["k"]
    D, I = flat_ip.flat_ip_search(g["q_txt"], g["txt"], k)
    assert np.array_equal(I, g["dpr_ids"]) and np.allclose(D, g["dpr_scores"], rtol=1e-6)
    # "L2norm,Flat": corpus normalised on add, queries normalised by KnowledgeBase.search_batch and again by FAISS
    D2, I2 = flat_ip.flat_ip_search(flat_ip.l2norm(g["q_img"]), flat_ip.l2norm(g["img"]), k, l2norm_queries=True)
    assert np.array_equal(I2, g["clip_ids"]) and np.allclose(D2, g["clip_scores"], rtol=1e-5, atol=1e-6)
    assert meta["do_L2norm"] == {"dpr": False, "clip": True}
    assert np.all(np.diff(g["dpr_scores"], axis=1) <= 0), "reference returns descending scores"
    assert g["dpr_ids"].dtype == np.int64 and g["dpr_scores"].dtype == np.float32


def test_c_oracle_equals_numpy_oracle():
    rng = np.random.default_rng(3)
    Q = rng.standard_normal((9, 48), dtype=np.float32)
    X = rng.standard_normal((3000, 48), dtype=np.float32)
    D, I = flat_ip.flat_ip_search(Q, X, 17)
    Dc, Ic = c_oracle.flat_ip(Q, X, 17)
    assert np.array_equal(I, Ic)
    assert np.allclose(D, Dc, rtol=1e-5, atol=1e-6)  # BLAS vs fixed-order fmaf: summation order only
    Df, If = flat_ip.flat_ip_search_fast(Q, X, 17)
    assert np.array_equal(I, If)


@pytest.mark.parametrize("dist", ["normal", "dpr_like", "unit"])
def test_oracle_against_independent_referee(dist):
    """The contraction + selection of both oracles against a referee that shares no code with them
    (tests/dropin.py:mips_referee — scikit-learn brute-force kNN in float64 over the MIPS -> L2 augmentation): same
    ids (no exact ties in this data: every rank-k gap is checked to exceed the fp32 accumulation error), scores
    within 1e-5 relative — the fp32 bar `north_star` states for FAISS IndexFlatIP."""
    from dropin import mips_referee
    rng = np.random.default_rng(17)
    n, d, nq, k = 20000, 96, 40, 100
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if dist == "dpr_like":   # a common mean direction: scores cluster like real DPR scores
        m = rng.standard_normal(d).astype(np.float32) * 0.8
        X, Q = (m + 0.35 * X).astype(np.float32), (m + 0.35 * Q).astype(np.float32)
    elif dist == "unit":     # CLIP / ArcFace regime
        X, Q = flat_ip.l2norm(X), flat_ip.l2norm(Q)
    Dr, Ir = mips_referee(Q, X, k + 1)
    gap = (Dr[:, :-1] - Dr[:, 1:]).min()
    Dr, Ir = Dr[:, :k], Ir[:, :k]
    scale = np.abs(Dr).max()
    for name, (D, I) in {"numpy": flat_ip.flat_ip_search(Q, X, k), "c": c_oracle.flat_ip(Q, X, k),
                         "fast": flat_ip.flat_ip_search_fast(Q, X, k)}.items():
        D, I = np.asarray(D), np.asarray(I)
        rel = np.abs(D - Dr) / np.maximum(np.abs(Dr), 1e-30)
        assert rel.max() <= 1e-5, (name, rel.max())
        differ = I != Ir
        if differ.any():  # only allowed between neighbours closer than the fp32 accumulation error
            assert gap < 1e-5 * scale and np.abs(D - Dr)[differ].max() <= 1e-5 * scale, (name, int(differ.sum()))
        else:
            assert np.array_equal(I, Ir), name


def test_c_oracle_is_fixed_order_fma():
    rng = np.random.default_rng(4)
    q = rng.standard_normal(37).astype(np.float32)
    x = rng.standard_normal((5, 37)).astype(np.float32)
    D, I = c_oracle.flat_ip(q[None], x, 5)
    for score, row in zip(D[0], I[0]):
        acc = 0.0  # float64 emulation of fmaf: exact product, one rounding per step
        for j in range(37):
            acc = float(np.float32(np.float64(q[j]) * np.float64(x[row, j]) + np.float64(np.float32(acc))))
        assert np.float32(acc) == score


def test_padding_ties_and_unselectable_values():
    X = np.array([[1, 0], [1, 0], [0, 1], [np.nan, 0], [-np.inf, 0]], dtype=np.float32)
    Q = np.array([[1, 0]], dtype=np.float32)
    for D, I in (flat_ip.flat_ip_search(Q, X, 6), c_oracle.flat_ip(Q, X, 6)):
        assert I[0].tolist()[:3] == [0, 1, 2], "ties: lower id first"
        assert I[0].tolist()[3:] == [-1, -1, -1], "NaN / -inf are never selected; padding id is -1"
        assert np.all(D[0][3:] == np.finfo(np.float32).min)
    D, I = flat_ip.flat_ip_search(np.zeros((0, 2), np.float32), X, 3)
    assert D.shape == (0, 3)
    with pytest.raises(ValueError):
        flat_ip.flat_ip_search(np.zeros(2, np.float32), X, 3)  # datasets/search.py:380-381


def test_l2norm_zero_row_is_nan_like_reference():
    out = flat_ip.l2norm(np.array([[3, 4], [0, 0]], np.float32))
    assert np.allclose(out[0], [0.6, 0.8]) and np.isnan(out[1]).all()
    assert np.allclose(c_oracle.l2norm(np.array([[3, 4]], np.float32)), [[0.6, 0.8]])


def test_bf16_rounding_matches_torch():
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 100
    assert np.array_equal(flat_ip.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).float().numpy())
    bits = c_oracle.to_bf16_bits(x)
    assert np.array_equal(bits, torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))


def test_merge_topk_is_topk_of_union():
    rng = np.random.default_rng(5)
    Q = rng.standard_normal((6, 16), dtype=np.float32)
    X = rng.standard_normal((900, 16), dtype=np.float32)
    D, I = flat_ip.flat_ip_search(Q, X, 20)
    parts = [flat_ip.flat_ip_search(Q, X[a:b], 20, id_offset=a) for a, b in ((0, 300), (300, 650), (650, 900))]
    Dm, Im = flat_ip.merge_topk([p[0] for p in parts], [p[1] for p in parts], 20)
    assert np.array_equal(I, Im) and np.array_equal(D, Dm)


# ---------------------------------------------------------------- fusion
def test_fusion_golden_default_minimum(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    mine = fuse_oracle.default_minimum(g["runs"])
    assert mine == g["after_defmin"], "restated default_minimum differs from the reference's output"
    assert all(len(r) == 0 for r in (mine[2][q] for q in ("q0", "q3")))  # runs without results stay empty


def test_fusion_golden_gzmuv(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    for src, key in ((g["after_defmin"], "gzmuv_after_defmin"), (g["runs"], "gzmuv_raw")):
        for run, ref in zip(src, g[key]):
            mine = fuse_oracle.gzmuv_norm(run)
            assert mine.keys() == ref.keys()
            for q in ref:
                assert mine[q].keys() == ref[q].keys()
                for d in ref[q]:
                    assert abs(mine[q][d] - ref[q][d]) <= 1e-12 * max(1.0, abs(ref[q][d]))


def test_survey_known_answer_gzmuv():
    out = fuse_oracle.gzmuv_norm({"q1": {"a": 3.0, "b": 1.0}, "q2": {"c": 5.0}})
    assert abs(out["q1"]["a"]) < 1e-12 and abs(out["q1"]["b"] + 1.224744871) < 1e-8 and abs(out["q2"]["c"] - 1.224744871) < 1e-8


def test_wsum_union_and_order():
    a = {"q": {"1": 1.0, "2": 0.5}}
    b = {"q": {"2": 1.0, "3": 0.25}}
    fused = fuse_oracle.wsum([a, b], [0.5, 0.5])
    assert list(fused["q"].items()) == [("2", 0.75), ("1", 0.5), ("3", 0.125)]
    z = fuse_oracle.zmuv_norm({"q": {"1": 2.0}})
    assert z["q"]["1"] == 0.0  # single result: std 0 -> denominator 1e-9


def test_fusion_oracle_against_independent_referees():
    """ranx is not installable here, so zmuv / min-max / wsum are restated from its published definitions (parity
    unpinned).  Narrowed with referees that share no code with oracle/fuse_oracle.py: scikit-learn's StandardScaler
    (population std) and MinMaxScaler per query for the two normalisations, a pandas outer-join for the weighted sum
    over the union of documents, and a stable numpy sort for the ranking."""
    import pandas as pd
    from sklearn.preprocessing import MinMaxScaler, StandardScaler

    rng = np.random.default_rng(11)
    runs = []
    for r in range(3):
        run = {}
        for q in range(7):
            docs = rng.choice(60, size=int(rng.integers(2, 25)), replace=False)
            run[q] = {int(d): float(s) for d, s in zip(docs, rng.normal(70 - 30 * r, 2 + r, size=len(docs)))}
        runs.append(run)
    for name, scaler in (("zmuv", StandardScaler), ("min-max", MinMaxScaler)):
        got = fuse_oracle.NORMS[name](runs[0])
        for q, results in runs[0].items():
            ref = scaler().fit_transform(np.array(list(results.values()), dtype=np.float64)[:, None])[:, 0]
            assert np.allclose(np.array([got[q][d] for d in results]), ref, rtol=1e-12, atol=1e-12), (name, q)
    weights = [0.5, 0.3, 0.2]
    fused = fuse_oracle.fuse(runs, weights, norm="zmuv", defmin=False)
    normed = [fuse_oracle.zmuv_norm(r) for r in runs]
    for q in range(7):
        frame = pd.concat([pd.Series(n[q], name=i, dtype=np.float64) for i, n in enumerate(normed)], axis=1).fillna(0.0)
        ref = (frame * weights).sum(axis=1)                      # a run that lacks a document contributes nothing
        ref = ref.sort_index().sort_values(ascending=False, kind="stable")   # descending score, ties by ascending doc id
        assert list(fused[q].keys()) == [int(d) for d in ref.index]
        assert np.allclose(np.array(list(fused[q].values())), ref.to_numpy(), rtol=1e-12, atol=1e-12)
    # default-minimum then gzmuv (the reference's own functions, pinned by the golden fixtures) keep the union complete
    filled = fuse_oracle.default_minimum(runs)
    for q in range(7):
        union = set().union(*[set(r[q]) for r in runs])
        assert all(set(f[q]) == union for f in filled)


def test_f32x3_split_is_exact_and_six_terms_are_fp32_faithful():
    """The VQ_F32X3 layout on the CPU: three bf16 terms reproduce every fp32 value exactly, and the six product
    terms the tensor cores accumulate miss the exact inner product by less than 2^-23 of sum |q_j x_j|."""
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.standard_normal(200_000).astype(np.float32),
                        (rng.standard_normal(50_000) * 1e-20).astype(np.float32), (rng.standard_normal(50_000) * 1e20).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 3.3895314e38, 1.1754944e-38, 65504.0, 1 + 2.0 ** -23], np.float32)])
    hi, mid, lo = flat_ip.split3_bf16(x)
    back = (hi + (mid + lo)).astype(np.float32)
    assert np.array_equal(back, x)  # value-exact (a zero's sign is restored from hi in join_rows_kernel)
    for t in (hi, mid, lo):
        assert np.array_equal(flat_ip.bf16_round(t), t)  # every term is a bf16 value
    Q = rng.standard_normal((16, 768)).astype(np.float32)
    X = rng.standard_normal((4000, 768)).astype(np.float32)
    exact = Q.astype(np.float64) @ X.astype(np.float64).T
    got = flat_ip.f32x3_scores(Q, X)
    bound = (np.abs(Q).astype(np.float64) @ np.abs(X).astype(np.float64).T) * 2.0 ** -23
    assert np.all(np.abs(got - exact) <= bound)
    assert float(np.max(np.abs(got - exact) / np.maximum(np.abs(exact), np.sqrt(768)))) < 2e-7
