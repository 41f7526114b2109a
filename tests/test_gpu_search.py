"""Parity of the CUDA path against the oracle — every call goes through the C ABI
(B200FlatIndex -> libviquae_b200.so).  Integer work (ids, fp32-mode scores) must be bit-exact;
16-bit modes are compared on the SAME quantised inputs within the tolerance BASELINE.json states:
scores within 2e-3 relative, top-k overlap >= 99.9 %, id differences only at (near-)ties."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, flat_ip

pytestmark = pytest.mark.gpu
REL_TOL_16 = 2e-3   # north star: bf16/fp16 scores within 2e-3 relative
MIN_OVERLAP = 0.999  # north star: top-100 overlap >= 99.9 %


def make(N, nq, d, dtype, seed=1, dist="normal"):
    rng = np.random.default_rng(seed)
    if dist == "dpr":  # DPR-like: common offset, scores cluster near 71 +- 2
        m = rng.standard_normal(d).astype(np.float32) * (8.0 / np.sqrt(d))
        X = m + 0.35 * rng.standard_normal((N, d), dtype=np.float32)
        Q = m * 8.9 + 0.35 * rng.standard_normal((nq, d), dtype=np.float32)
    else:
        X = rng.standard_normal((N, d), dtype=np.float32)
        Q = rng.standard_normal((nq, d), dtype=np.float32)
    if dtype == "bf16":
        X, Q = flat_ip.bf16_round(X), flat_ip.bf16_round(Q)
    elif dtype == "f16":
        X, Q = flat_ip.fp16_round(X), flat_ip.fp16_round(Q)
    return X, Q


def engine(X, dtype, variant=None, **kw):
    from viquae_b200.index import B200FlatIndex
    idx = B200FlatIndex(d=X.shape[1], dtype=dtype, device=0, variant=variant, **kw)
    idx.add(X)
    return idx


def check16(idx, X, Q, k):
    res = idx.search_batch(Q, k)
    D, I = np.asarray(res.total_scores), np.asarray(res.total_indices)
    Dr, Ir = flat_ip.flat_ip_search(Q, X, k)
    cmp = flat_ip.compare_results(Dr, Ir, D, I, rel_tol=REL_TOL_16)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] >= MIN_OVERLAP and cmp["max_rel_err"] <= REL_TOL_16, cmp
    assert np.all(np.diff(D, axis=1) <= 0), "scores must be descending"
    assert D.dtype == np.float32 and I.dtype.kind == "i"
    return D, I


@pytest.mark.parametrize("N,nq,d,k", [(5000, 70, 768, 100), (3001, 130, 100, 10), (20000, 3, 512, 128), (257, 1, 8, 5)])
def test_fp32_mode_is_bit_exact(N, nq, d, k):
    X, Q = make(N, nq, d, "f32")
    idx = engine(X, "f32")
    res = idx.search_batch(Q, k)
    Dr, Ir = c_oracle.flat_ip(Q, X, k)
    assert np.array_equal(np.asarray(res.total_indices), Ir)
    assert np.array_equal(np.asarray(res.total_scores).view(np.uint32), Dr.view(np.uint32)), "fixed reduction order: scores bit-exact"


@pytest.mark.parametrize("d", [100, 33, 768])
def test_fp32_tiled_storage_ragged_adds_reconstruct_and_l2norm(d, tmp_path):
    """The fp32 exact mode keeps its rows as k-major tiles of 128 rows (what K1's TMA boxes read): appends that start and
    end inside a tile, growth across reallocations, d not a multiple of the 32-column k block, reconstruct / save / load
    giving back the exact bits, device-resident sources, and the ingest-side L2 normalisation producing the same bits as
    the row-major prep kernel behind vq_prep_rows."""
    from viquae_b200 import _lib
    from viquae_b200.index import B200FlatIndex
    X, Q = make(1500, 37, d, "f32", seed=5)
    X[1153:] = flat_ip.bf16_round(X[1153:])      # the last append arrives as a bf16 tensor (exact in fp32)
    idx = B200FlatIndex(d=d, dtype="f32")
    cuts = [0, 77, 78, 377, 1024, 1153, 1500]   # ragged appends: inside a tile, exactly to a tile edge, across several tiles
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a == 1153:
            idx.add(torch.from_numpy(X[a:b]).cuda().to(torch.bfloat16))   # 16-bit device source, widened on ingest
        elif (a // 100) % 2:
            idx.add(torch.from_numpy(X[a:b]).cuda())     # device-resident source
        else:
            idx.add(X[a:b])                               # host source through the pinned pipeline
    assert len(idx) == 1500
    assert np.array_equal(idx.reconstruct_n(0, None).view(np.uint32), X.view(np.uint32))
    assert np.array_equal(idx.reconstruct_n(120, 20).view(np.uint32), X[120:140].view(np.uint32))   # a window across a tile edge
    for k in (1, 10):
        res = idx.search_batch(Q, k)
        Dr, Ir = c_oracle.flat_ip(Q, X, k)
        assert np.array_equal(np.asarray(res.total_indices), Ir)
        assert np.array_equal(np.asarray(res.total_scores).view(np.uint32), Dr.view(np.uint32))
    idx.save(tmp_path / "f32.index")
    idx2 = B200FlatIndex.load(tmp_path / "f32.index")
    r1, r2 = idx.search_batch(Q, 10), idx2.search_batch(Q, 10)
    assert np.array_equal(r1.total_indices, r2.total_indices) and np.array_equal(r1.total_scores, r2.total_scores)
    idx.close(); idx2.close()
    # L2norm on ingest: same bits as the row-major prep kernel (one zero row stays zero, FAISS semantics)
    Xz = X[:300].copy(); Xz[17] = 0.0
    idn = B200FlatIndex(d=d, dtype="f32", string_factory="L2norm,Flat")
    idn.add(Xz)
    src = torch.from_numpy(Xz).cuda()
    ref = torch.empty_like(src)
    _lib.check(_lib.load().vq_prep_rows(src.data_ptr(), _lib.F32, ref.data_ptr(), _lib.F32, 300, d, 2, None))
    torch.cuda.synchronize()
    assert np.array_equal(idn.reconstruct_n(0, None).view(np.uint32), ref.cpu().numpy().view(np.uint32))
    assert not np.any(idn.reconstruct_n(17, 1))
    idn.close()


@pytest.mark.parametrize("variant,N,nq,d,k,dtype", [
    ("tc_cg1", 5000, 70, 768, 100, "bf16"),
    ("tc_cg1", 70001, 100, 768, 100, "bf16"),
    ("tc_cg2", 5000, 300, 768, 100, "bf16"),
    ("tc_cg2", 70000, 1250, 768, 100, "bf16"),       # validation-size batch (BASELINE configs[0] queries)
    ("tc_cg2", 300000, 300, 768, 100, "bf16"),       # multi-pass threshold priming
    ("tc_cg1", 300100, 100, 512, 10, "f16"),         # CLIP/ArcFace dimension, fp16
    ("tc_cg2", 600000, 130, 64, 128, "bf16"),        # k at the supported maximum
    ("tc_cg1", 9000, 33, 104, 7, "bf16"),            # d not a multiple of 64 (row stride padded to 8)
    (None, 20000, 1, 768, 100, "bf16"),               # latency regime: one query (3-pass schedule, CTA-wide merge sort)
    (None, 20000, 20, 768, 100, "bf16"),
    (None, 5000, 32, 512, 10, "f16"),
    (None, 400000, 8, 768, 100, "bf16"),              # several priming passes at a small batch
    (None, 40000, 256, 768, 100, "bf16"),            # the reference's call granularity: 256 queries per call
    (None, 30000, 140, 2048, 100, "bf16"),           # imagenet-RN50 dimension (config_test.json:17-40)
    (None, 30000, 40, 1024, 100, "f16"),             # clip-RN50 dimension
])
def test_16bit_modes_match_oracle_on_same_quantised_inputs(variant, N, nq, d, k, dtype):
    X, Q = make(N, nq, d, dtype)
    check16(engine(X, dtype, variant), X, Q, k)


def test_dpr_like_scores_and_l2norm_factory():
    X, Q = make(30000, 64, 768, "bf16", dist="dpr")
    check16(engine(X, "bf16"), X, Q, 100)
    # "L2norm,Flat": rows normalised on ingest, queries on search (cosine); compare in fp32 mode against the oracle
    X, Q = make(4000, 40, 512, "f32", seed=3)
    idx = engine(X, "f32", string_factory="L2norm,Flat")
    res = idx.search_batch(Q, 20)
    Dr, Ir = flat_ip.flat_ip_search(flat_ip.l2norm(Q), flat_ip.l2norm(X), 20)
    cmp = flat_ip.compare_results(Dr, Ir, res.total_scores, res.total_indices, rel_tol=1e-5, abs_tol=1e-6)
    assert cmp["real_mismatches"] == 0 and cmp["max_rel_err"] < 1e-4 and np.abs(np.asarray(res.total_scores)).max() <= 1.0001, cmp


def test_edge_cases():
    from viquae_b200.index import B200FlatIndex
    X, Q = make(50, 4, 64, "bf16")
    idx = engine(X, "bf16")
    res = idx.search_batch(Q, 100)          # N < k: padded with (-FLT_MAX, -1)
    I, D = np.asarray(res.total_indices), np.asarray(res.total_scores)
    assert (I[:, 50:] == -1).all() and (D[:, 50:] == np.finfo(np.float32).min).all() and (I[:, :50] >= 0).all()
    Dr, Ir = flat_ip.flat_ip_search(Q, X, 100)
    assert flat_ip.compare_results(Dr, Ir, D, I, rel_tol=REL_TOL_16)["real_mismatches"] == 0
    assert idx.search_batch(np.zeros((0, 64), np.float32), 5).total_scores.shape == (0, 5)   # empty batch
    with pytest.raises(ValueError):
        idx.search_batch(np.zeros(64, np.float32), 5)                                        # datasets/search.py:380-381
    with pytest.raises(ValueError):
        idx.search_batch(np.zeros((2, 63), np.float32), 5)
    from viquae_b200._lib import EngineError
    with pytest.raises(EngineError):
        idx.search_batch(Q, 2049)                                                            # k > 2048: refused, not approximated
    one = idx.search(Q[0], 7)                                                                 # FaissIndex.search: one vector
    assert len(one.scores) == 7 and np.array_equal(np.asarray(one.indices), I[0, :7])
    empty = B200FlatIndex(d=64, dtype="bf16")
    r = empty.search_batch(Q, 3)
    assert (np.asarray(r.total_indices) == -1).all()
    # exact ties: duplicated rows -> lower id first, on every kernel family
    Xd = np.repeat(X[:10], 3, axis=0)
    for dtype, variant in (("bf16", "tc_cg1"), ("bf16", "tc_cg2"), ("f32", None)):
        r = engine(Xd, dtype, variant).search_batch(Q, 30)
        ids = np.asarray(r.total_indices)
        s = np.asarray(r.total_scores)
        for row_s, row_i in zip(s, ids):
            for a in range(0, 30, 3):
                assert row_s[a] == row_s[a + 1] == row_s[a + 2] and row_i[a] < row_i[a + 1] < row_i[a + 2]
    # NaN rows are never selected (FAISS: a candidate enters only if strictly greater)
    Xn = X.copy(); Xn[7] = np.nan
    r = engine(Xn, "f32").search_batch(Q, 50)
    assert 7 not in np.asarray(r.total_indices) and (np.asarray(r.total_indices)[:, 49] == -1).all()


def test_incremental_add_growth_save_load(tmp_path):
    X, Q = make(7000, 16, 96, "bf16")
    from viquae_b200.index import B200FlatIndex
    idx = B200FlatIndex(d=96, dtype="bf16")
    for a in range(0, 7000, 1000):          # the reference adds 1000-row batches (datasets/search.py:311-313)
        idx.add(X[a:a + 1000])
    assert len(idx) == 7000
    D, I = check16(idx, X, Q, 10)
    idx.add(torch.from_numpy(X[:100]).cuda().to(torch.bfloat16))   # device-resident ingest
    assert len(idx) == 7100
    idx.save(tmp_path / "kb.index")
    idx2 = B200FlatIndex.load(tmp_path / "kb.index")
    r1, r2 = idx.search_batch(Q, 10), idx2.search_batch(Q, 10)
    assert np.array_equal(r1.total_indices, r2.total_indices) and np.array_equal(r1.total_scores, r2.total_scores)
    assert np.array_equal(idx.reconstruct_n(5, 3), X[5:8])


def test_device_resident_search_and_cross_shard_merge():
    from viquae_b200.index import B200FlatIndex, merge_topk
    X, Q = make(9000, 40, 128, "bf16")
    full = engine(X, "bf16")
    s_full, i_full = full.search_tensors(torch.from_numpy(Q).cuda(), 50)
    parts = []
    for a, b in ((0, 3000), (3000, 6500), (6500, 9000)):
        sh = B200FlatIndex(d=128, dtype="bf16", id_offset=a)
        sh.add(X[a:b])
        parts.append(sh.search_tensors(torch.from_numpy(Q).cuda().to(torch.bfloat16), 50))
    s_m, i_m = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    cmp = flat_ip.compare_results(s_full.cpu().numpy(), i_full.cpu().numpy(), s_m.cpu().numpy(), i_m.cpu().numpy(), rel_tol=1e-5)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] == 1.0, cmp
    Dm, Im = flat_ip.merge_topk([p[0].cpu().numpy() for p in parts], [p[1].cpu().numpy() for p in parts], 50)
    assert np.array_equal(Im, i_m.cpu().numpy()) and np.array_equal(Dm, s_m.cpu().numpy())


def test_reference_golden_through_knowledgebase_and_searcher(golden_dir, tmp_path):
    """The toy KB of tests/golden (outputs of the reference's own KnowledgeBase/Searcher) through OUR
    KnowledgeBase/Searcher with kind=B200, fp32 storage: same ids, same str doc ids, same runs."""
    from viquae_b200.search import KnowledgeBase, Searcher
    g = np.load(os.path.join(golden_dir, "search_golden.npz"))
    meta = json.load(open(os.path.join(golden_dir, "search_golden.json")))
    k = meta["k"]
    kb_data = {"text_emb": g["txt"], "image_emb": g["img"]}
    map_path = tmp_path / "a2p.json"
    json.dump(meta["mapping"], open(map_path, "wt"))
    qrels_path = tmp_path / "qrels.json"
    json.dump({}, open(qrels_path, "wt"))
    index_kwargs = {"dpr": {"column": "text_emb", "kind": "B200", "key": "q_text", "string_factory": "Flat", "metric_type": 0, "dtype": "f32"},
                    "clip": {"column": "image_emb", "kind": "B200", "key": "q_image", "string_factory": "L2norm,Flat", "metric_type": 0, "dtype": "f32"}}

    class DictDS(dict):
        def __len__(self):
            return len(next(iter(self.values())))
    kb = KnowledgeBase(dataset=DictDS(kb_data), index_mapping_path=str(map_path), index_kwargs=index_kwargs)
    assert {n: ix.do_L2norm for n, ix in kb.indexes.items()} == meta["do_L2norm"]
    s1, i1 = kb.search_batch("dpr", g["q_txt"], k=k)
    assert np.array_equal(np.asarray(i1), g["dpr_ids"]) and np.allclose(s1, g["dpr_scores"], rtol=1e-5)
    s2, i2 = kb.search_batch("clip", g["q_img"], k=k)
    assert np.array_equal(np.asarray(i2), g["clip_ids"]) and np.allclose(s2, g["clip_scores"], rtol=1e-4, atol=1e-6)
    q_img_list = [None if m else g["q_img"][j].tolist() for j, m in enumerate(g["none_mask"])]
    s3, i3 = kb.search_batch_if_not_None("clip", q_img_list, k=k)
    assert [len(x) for x in i3] == meta["if_not_none_lengths"]
    searcher = Searcher.__new__(Searcher)
    searcher.k, searcher.kbs, searcher.runs = k, {"kb": kb}, {"dpr": {}, "clip": {}}
    searcher.qrels, searcher.qnonrels, searcher.reference_kb, searcher.find_relevant = {}, {}, None, None
    nq = len(g["q_txt"])
    searcher({"id": [f"q{j}" for j in range(nq)], "q_text": g["q_txt"].tolist(), "q_image": q_img_list})
    for name in ("dpr", "clip"):
        for q, ref in meta["runs"][name].items():
            mine = searcher.runs[name][q]
            assert list(mine) == list(ref), (name, q)
            assert np.allclose([float(v) for v in mine.values()], list(ref.values()), rtol=1e-4, atol=1e-6)


def test_large_corpus_properties():
    """Size-independent properties at a size the oracle cannot check in seconds: 2M x 768 bf16 generated on the
    device.  (i) returned scores equal recomputed dot products of the returned rows, (ii) descending order,
    (iii) sharding invariance: top-k of the whole == merge of the top-k of two halves, (iv) determinism."""
    from viquae_b200.index import B200FlatIndex, merge_topk
    N, d, nq, k = 2_000_000, 768, 1250, 100
    g = torch.Generator(device="cuda"); g.manual_seed(123)
    X = torch.randn((N, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    Q = torch.randn((nq, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    idx = B200FlatIndex(d=d, dtype="bf16", capacity=N); idx.add(X)
    s, i = idx.search_tensors(Q, k)
    assert (i >= 0).all() and (i < N).all() and (s[:, 1:] <= s[:, :-1]).all()
    rows = X[i.reshape(-1)].float().reshape(nq, k, d)
    recomputed = torch.einsum("qkd,qd->qk", rows, Q.float())
    assert torch.allclose(recomputed, s, rtol=REL_TOL_16, atol=1e-3)
    # a random sample of rows never beats the k-th returned score
    sample = torch.randint(0, N, (4096,), device="cuda")
    assert ((Q.float() @ X[sample].float().T).max(dim=1).values <= s[:, 0] + 1e-3).all()
    s2, i2 = idx.search_tensors(Q, k)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    halves = []
    for a, b in ((0, N // 2), (N // 2, N)):
        sh = B200FlatIndex(d=d, dtype="bf16", capacity=b - a, id_offset=a); sh.add(X[a:b])
        halves.append(sh.search_tensors(Q, k)); sh.close()
    s_m, i_m = merge_topk(torch.stack([h[0] for h in halves]), torch.stack([h[1] for h in halves]))
    cmp = flat_ip.compare_results(s.cpu().numpy(), i.cpu().numpy(), s_m.cpu().numpy(), i_m.cpu().numpy(), rel_tol=1e-4)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] >= 0.9999, cmp
    # small batch over the same corpus agrees with the large batch (different kernel variant / pass schedule)
    s3, i3 = idx.search_tensors(Q[:5], k)
    cmp = flat_ip.compare_results(s[:5].cpu().numpy(), i[:5].cpu().numpy(), s3.cpu().numpy(), i3.cpu().numpy(), rel_tol=1e-4)
    assert cmp["real_mismatches"] == 0, cmp


def test_hf_dataset_dispatch_and_dataset_search_with_fusion(tmp_path):
    """HF datasets integration: the engine registered in Dataset._indexes serves Dataset.search_batch
    (datasets/search.py:713-733), and dataset_search() maps a query dataset, writes the runs and fuses them."""
    datasets = pytest.importorskip("datasets")
    from viquae_b200.search import KnowledgeBase, dataset_search
    rng = np.random.default_rng(9)
    n, d1, d2, nq, k = 3000, 64, 32, 40, 20
    txt = rng.standard_normal((n, d1)).astype(np.float32)
    img = rng.standard_normal((n, d2)).astype(np.float32)
    kb_ds = datasets.Dataset.from_dict({"text_emb": txt.tolist(), "image_emb": img.tolist()})
    kb_path = tmp_path / "kb"
    kb_ds.save_to_disk(str(kb_path))
    kb = KnowledgeBase(str(kb_path), index_kwargs={"dpr": {"column": "text_emb", "kind": "B200", "key": "q_text", "string_factory": "Flat",
                                                            "metric_type": 0, "dtype": "f32"}})
    q = rng.standard_normal((nq, d1)).astype(np.float32)
    res = kb.dataset.search_batch("dpr", q, k)              # HF entry point -> B200FlatIndex.search_batch
    Dr, Ir = c_oracle.flat_ip(q, txt, k)
    assert np.array_equal(np.asarray(res.total_indices), Ir)
    # full pipeline: two indexes, qrels from the oracle's top-1 of the text index, fusion applied with fixed weights
    qi = rng.standard_normal((nq, d2)).astype(np.float32)
    queries = datasets.Dataset.from_dict({"id": [f"q{j}" for j in range(nq)], "q_text": q.tolist(), "q_image": qi.tolist(),
                                          "output": [{"original_answer": "", "answer": []}] * nq})
    qrels = {f"q{j}": {str(int(Ir[j, 0])): 1} for j in range(nq)}
    json.dump(qrels, open(tmp_path / "qrels.json", "wt"))
    out = tmp_path / "metrics"
    searcher = dataset_search(
        queries, k=k, metric_save_path=out, map_kwargs={"batch_size": 16},
        kb_kwargs={str(kb_path): {"index_kwargs": {
            "dpr": {"column": "text_emb", "kind": "B200", "key": "q_text", "string_factory": "Flat", "metric_type": 0, "dtype": "f32"},
            "clip": {"column": "image_emb", "key": "q_image", "string_factory": "L2norm,Flat", "metric_type": 0}}}},
        qrels=str(tmp_path / "qrels.json"),
        fusion_kwargs={"subcommand": "test", "subcommand_kwargs": {"best_params": {"weights": [0.8, 0.2]}}, "norm": "zmuv", "defmin": True})
    assert (out / "dpr.json").exists() and (out / "clip.json").exists() and (out / "test_run.json").exists() and (out / "metrics.json").exists()
    metrics = json.load(open(out / "metrics.json"))
    assert metrics["dpr"]["mrr@100"] == 1.0 and metrics["dpr"]["precision@1"] == 1.0
    fused = json.load(open(out / "test_run.json"))
    from oracle import fuse_oracle
    ref = fuse_oracle.fuse([searcher.runs["dpr"].run, searcher.runs["clip"].run], [0.8, 0.2], norm="zmuv", defmin=True)
    for qid in ("q0", "q7", "q39"):
        assert set(fused[qid]) == set(ref[qid])
        assert np.allclose([fused[qid][d_] for d_ in ref[qid]], list(ref[qid].values()), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("many2one", [None, "max"])
@pytest.mark.parametrize("score_range", ["dpr", "cosine"])
def test_gpu_postprocessing_equals_reference_loop(many2one, score_range):
    """K6 (vq_expand_hits) vs the Python restatement of Searcher.__call__'s loop (search.py:413-440), including
    overlapping mappings (dict overwrite / max semantics), early stop at k and padded hits.  "cosine": scores in
    [0, 1] as the `L2norm,Flat` indexes return them, where the reference's 1e-8 per-passage penalty is visible in
    float32 and must be reproduced bit for bit."""
    from viquae_b200.postprocess import CSRMapping, expand_hits
    from viquae_b200.search import Searcher
    rng = np.random.default_rng(2)
    n_rows, nq, k = 500, 64, 20
    mapping = {r: sorted(rng.choice(300, size=int(rng.integers(0, 6)), replace=False).tolist()) for r in range(n_rows)}
    raw = rng.normal(70, 2, size=(nq, k)) if score_range == "dpr" else rng.uniform(0.0, 1.0, size=(nq, k))
    scores = np.sort(raw.astype(np.float32), axis=1)[:, ::-1].copy()
    if score_range == "cosine":
        scores[3, 4] = scores[3, 3]  # an exact tie between two articles
    ids = np.stack([rng.choice(n_rows, size=k, replace=False) for _ in range(nq)]).astype(np.int64)
    ids[5, 7:] = -1  # a row with fewer than k hits
    scores[5, 7:] = np.finfo(np.float32).min

    class KB:
        index_mapping = mapping
    KB.many2one = many2one
    s = Searcher.__new__(Searcher)
    s.k = k
    out_s, out_i, out_c = expand_hits(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), CSRMapping(mapping, n_rows=n_rows), many2one, k)
    out_s, out_i, out_c = out_s.cpu().numpy(), out_i.cpu().numpy(), out_c.cpu().numpy()
    for q in range(nq):
        ref = {}
        valid = ids[q] >= 0
        s._record(KB, ref, scores[q][valid], ids[q][valid])
        n = out_c[q]
        assert [str(d) for d in out_i[q, :n]] == list(ref), q          # same documents in the same (insertion) order
        assert np.array_equal(out_s[q, :n].astype(np.float32), np.array(list(ref.values()), dtype=np.float32))
        assert (out_i[q, n:] == -1).all()
    # identity mapping: the run is the hit list
    out_s, out_i, out_c = expand_hits(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), None, None, k)
    assert np.array_equal(out_i.cpu().numpy()[0], ids[0]) and out_c.cpu().numpy()[5] == 7


def test_full_size_corpus_properties():
    """BASELINE configs[1] at full size (12M x 768 bf16 on one GPU, generated on the device): the oracle cannot
    run here, so size-independent properties: recomputed scores, ordering, determinism, sharding invariance
    against two 6M-row halves merged by K4, and agreement between the 256-query and 1-query schedules."""
    from viquae_b200.index import B200FlatIndex, merge_topk
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs > 60 GB of HBM")
    N, d, nq, k, chunk = 12_000_000, 768, 256, 100, 1_000_000
    idx = B200FlatIndex(d=d, dtype="bf16", capacity=N)
    halves = [B200FlatIndex(d=d, dtype="bf16", capacity=N // 2, id_offset=0), B200FlatIndex(d=d, dtype="bf16", capacity=N // 2, id_offset=N // 2)]
    for c in range(N // chunk):
        g = torch.Generator(device="cuda"); g.manual_seed(20240117 + c)
        x = torch.randn((chunk, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
        idx.add(x)
        halves[0 if c < 6 else 1].add(x)
        if c == 7:
            keep = x[:1000].clone()   # rows 7,000,000 .. 7,000,999 for the recomputation check
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    Q = torch.randn((nq, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    Q[:8] = keep[:8] * 4   # plant queries whose best match is a known row
    s, i = idx.search_tensors(Q, k)
    assert (i[:8, 0] == torch.arange(7_000_000, 7_000_008, device="cuda")).all()
    assert torch.allclose(s[:8, 0], (keep[:8].float() * Q[:8].float()).sum(-1), rtol=2e-3)
    assert (i >= 0).all() and (i < N).all() and (s[:, 1:] <= s[:, :-1]).all()
    s2, i2 = idx.search_tensors(Q, k)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    parts = [h.search_tensors(Q, k) for h in halves]
    s_m, i_m = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    cmp = flat_ip.compare_results(s.cpu().numpy(), i.cpu().numpy(), s_m.cpu().numpy(), i_m.cpu().numpy(), rel_tol=1e-4)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] >= 0.9999, cmp
    s1, i1 = idx.search_tensors(Q[:1], k)
    cmp = flat_ip.compare_results(s[:1].cpu().numpy(), i[:1].cpu().numpy(), s1.cpu().numpy(), i1.cpu().numpy(), rel_tol=1e-4)
    assert cmp["real_mismatches"] == 0, cmp


def test_packed_merge_equals_plain_merge():
    from viquae_b200.index import merge_topk, merge_topk_packed, packed_result_buffer
    rng = np.random.default_rng(4)
    G, nq, k = 3, 37, 100
    s = np.sort(rng.standard_normal((G, nq, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    i = np.stack([rng.permutation(100000)[:nq * k].reshape(nq, k) + g * 100000 for g in range(G)]).astype(np.int64)
    i[1, 5, 60:] = -1
    buf, s_v, i_v = packed_result_buffer(nq, k, "cuda:0", n_lists=G)
    for g in range(G):
        s_v[g].copy_(torch.from_numpy(s[g]))
        i_v[g].copy_(torch.from_numpy(i[g]))
    a = merge_topk_packed(buf, nq, k)
    b = merge_topk(torch.from_numpy(s).cuda(), torch.from_numpy(i).cuda())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    Dm, Im = flat_ip.merge_topk(list(s), list(i), k)
    assert np.array_equal(Im, a[1].cpu().numpy()) and np.array_equal(Dm, a[0].cpu().numpy())


def test_randomized_small_shapes_all_paths():
    """Seeded fuzz over awkward shapes (d = 1.., N = 1.., k = 1..128, nq = 1..300) on every kernel family."""
    rng = np.random.default_rng(2024)
    for trial in range(36):
        d = int(rng.choice([1, 2, 3, 7, 8, 9, 31, 63, 64, 65, 100, 129, 200, 384]))
        N = int(rng.choice([1, 2, 5, 100, 255, 256, 257, 511, 1000, 4097]))
        nq = int(rng.choice([1, 2, 31, 33, 127, 128, 129, 257, 300]))
        k = int(rng.choice([1, 2, 5, 64, 100, 127, 128]))
        dtype, variant = [("bf16", None), ("bf16", "tc_cg1"), ("bf16", "tc_cg2"), ("f16", None), ("f32", None), ("f16", "tc_cg1")][trial % 6]
        X, Q = make(N, nq, d, dtype, seed=100 + trial)
        idx = engine(X, dtype, variant)
        res = idx.search_batch(Q, k)
        D, I = np.asarray(res.total_scores), np.asarray(res.total_indices)
        if dtype == "f32":
            Dr, Ir = c_oracle.flat_ip(Q, X, k)
            assert np.array_equal(I, Ir) and np.array_equal(D.view(np.uint32), Dr.view(np.uint32)), (trial, d, N, nq, k)
        else:
            Dr, Ir = flat_ip.flat_ip_search(Q, X, k)
            cmp = flat_ip.compare_results(Dr, Ir, D, I, rel_tol=REL_TOL_16, abs_tol=1e-4)
            assert cmp["real_mismatches"] == 0 and (I >= 0).sum() == (Ir >= 0).sum(), (trial, d, N, nq, k, dtype, variant, cmp)
        idx.close()


def test_fp32_exact_mode_fuzz_multi_pass_shapes():
    """Seeded fuzz of the fp32 exact mode alone (TMA ring + k-major tiles): shapes with one or several priming passes, more
    work items than CTAs and fewer, several query tiles, d around the 32-column k block — ids and score bits equal the
    fixed-order C oracle every time."""
    rng = np.random.default_rng(77)
    shapes = [(40000, 300, 96, 100), (70001, 5, 33, 10), (20000, 641, 64, 7), (129, 129, 32, 128), (38000, 128, 31, 1),
              (128, 1, 1, 1), (50000, 257, 40, 50), (9000, 700, 160, 20)]
    for trial in range(10):
        shapes.append((int(rng.integers(1, 30000)), int(rng.integers(1, 400)), int(rng.integers(1, 130)), int(rng.integers(1, 129))))
    for trial, (N, nq, d, k) in enumerate(shapes):
        X, Q = make(N, nq, d, "f32", seed=500 + trial)
        idx = engine(X, "f32")
        res = idx.search_batch(Q, k)
        Dr, Ir = c_oracle.flat_ip(Q, X, k)
        assert np.array_equal(np.asarray(res.total_indices), Ir), (trial, N, nq, d, k)
        assert np.array_equal(np.asarray(res.total_scores).view(np.uint32), Dr.view(np.uint32)), (trial, N, nq, d, k)
        idx.close()


def test_knowledgebase_save_path_then_load(tmp_path):
    """`save_path` / `load` of the index config (meerqat/ir/search.py:234-248), through KnowledgeBase."""
    from viquae_b200.search import KnowledgeBase
    rng = np.random.default_rng(12)
    emb = rng.standard_normal((800, 40)).astype(np.float32)
    q = rng.standard_normal((9, 40)).astype(np.float32)

    class DictDS(dict):
        def __len__(self):
            return len(next(iter(self.values())))
    path = str(tmp_path / "clip.index")
    kb1 = KnowledgeBase(dataset=DictDS({"e": emb}), index_kwargs={"clip": {"column": "e", "key": "q", "string_factory": "L2norm,Flat",
                                                                            "metric_type": 0, "save_path": path}})
    kb2 = KnowledgeBase(dataset=DictDS({"e": emb}), index_kwargs={"clip": {"column": "e", "key": "q", "string_factory": "L2norm,Flat",
                                                                            "metric_type": 0, "load": True, "file": path}})
    assert kb2.indexes["clip"].do_L2norm is True
    s1, i1 = kb1.search_batch("clip", q, k=10)
    s2, i2 = kb2.search_batch("clip", q, k=10)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    with pytest.raises(ValueError, match="Index size should match"):
        KnowledgeBase(dataset=DictDS({"e": emb[:10]}), index_kwargs={"clip": {"column": "e", "key": "q", "string_factory": "L2norm,Flat",
                                                                              "metric_type": 0, "load": True, "file": path}})


def test_faiss_flat_file_interchange(tmp_path):
    """An index written in FAISS's "Flat" / "L2norm,Flat" file layout (what the reference's `save_path` leaves on
    disk, meerqat/ir/search.py:246-248) loads through `load: true` and searches like the index it came from."""
    from viquae_b200 import faiss_io
    from viquae_b200.index import B200FlatIndex
    from viquae_b200.search import KnowledgeBase
    rng = np.random.default_rng(21)
    emb = rng.standard_normal((1500, 48)).astype(np.float32)
    q = rng.standard_normal((17, 48)).astype(np.float32)
    for factory in ("Flat", "L2norm,Flat"):
        src = B200FlatIndex(dtype="f32", string_factory=factory)
        src.add(emb)
        path = tmp_path / (factory.replace(",", "_") + ".faiss")
        faiss_io.save_faiss_flat(src, path)
        ff = faiss_io.read_flat_index(path)
        assert ff.string_factory == factory and ff.rows.shape == emb.shape
        if factory == "Flat":
            assert np.array_equal(ff.rows, emb)
        else:
            np.testing.assert_allclose(np.linalg.norm(ff.rows, axis=1), 1.0, rtol=1e-6)

        class DictDS(dict):
            def __len__(self):
                return len(next(iter(self.values())))
        kb = KnowledgeBase(dataset=DictDS({"e": emb}), index_kwargs={"x": {"column": "e", "key": "q", "string_factory": factory, "metric_type": 0,
                                                                           "dtype": "f32", "load": True, "file": str(path)}})
        assert kb.indexes["x"].do_L2norm is (factory != "Flat")
        s1, i1 = src.search_batch(q, k=20)
        s2, i2 = kb.search_batch("x", q, k=20)
        assert np.array_equal(np.asarray(i1), np.asarray(i2)) and np.array_equal(np.asarray(s1), np.asarray(s2))
    faiss_io.write_flat_index(tmp_path / "l2.faiss", emb, metric_type=1)
    with pytest.raises(ValueError, match="inner-product only"):
        B200FlatIndex.load(tmp_path / "l2.faiss")
    # What the reference leaves on disk when `device` is set (search.py:237-248): the column pre-normalised in Python,
    # "L2norm" stripped from the factory -> a PLAIN Flat file; the config still says "L2norm,Flat", and that flag (not
    # the file's) decides that queries get normalised.
    unit = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    faiss_io.write_flat_index(tmp_path / "stripped.faiss", unit, l2norm=False)
    kb = KnowledgeBase(dataset=DictDS({"e": emb}), index_kwargs={"x": {"column": "e", "key": "q", "string_factory": "L2norm,Flat", "metric_type": 0,
                                                                       "dtype": "f32", "load": True, "file": str(tmp_path / "stripped.faiss")}})
    assert kb.indexes["x"].do_L2norm is True and kb._engines["x"].do_L2norm is True
    s3, i3 = kb.search_batch("x", q * 7.5, k=20)   # unnormalised queries: scores must come back as cosines
    Dr, Ir = c_oracle.flat_ip(flat_ip.l2norm(q * 7.5), unit.astype(np.float32), 20)
    assert np.array_equal(np.asarray(i3), Ir) and np.allclose(s3, Dr, rtol=1e-5, atol=1e-6) and np.abs(s3).max() <= 1.0 + 1e-5


def test_zero_norm_corpus_row_keeps_faiss_semantics():
    """`"L2norm,Flat"` on the CPU path is FAISS's NormalizationTransform, which leaves a zero vector untouched: the row
    scores 0 and stays retrievable.  Only the query side follows numpy's q / ||q|| (zero query -> NaN scores, nothing
    selected), meerqat/ir/search.py:43-46."""
    from viquae_b200.index import B200FlatIndex
    rng = np.random.default_rng(4)
    emb = rng.standard_normal((300, 32)).astype(np.float32)
    emb[17] = 0.0
    for dtype in ("f32", "bf16", "f32x3"):
        idx = B200FlatIndex(d=32, dtype=dtype, string_factory="L2norm,Flat")
        idx.add(emb)
        rows = idx.reconstruct_n(17, 1, "f32")
        assert np.array_equal(rows, np.zeros((1, 32), np.float32)), dtype
        q = -emb[[3]]  # every cosine but row 17's is what it is; ask for all rows so that the zero row must appear
        res = idx.search_batch(q, k=300)
        ids = np.asarray(res.total_indices)[0]
        assert 17 in ids and np.asarray(res.total_scores)[0][list(ids).index(17)] == 0.0, dtype
        idx.close()


@pytest.mark.parametrize("dtype,N,nq,d,k", [("bf16", 30000, 140, 256, 129), ("bf16", 50000, 40, 768, 1000), ("f16", 3000, 9, 64, 300),
                                            ("f32", 6000, 33, 96, 500), ("bf16", 700, 5, 32, 2048), ("bf16", 300000, 260, 128, 257)])
def test_large_k_banded_rounds(dtype, N, nq, d, k):
    """k > 128: ceil(k/128) banded rounds, each taking what ranks strictly after the previous round's last hit.
    Exactness and tie order must be those of a single top-k (N < k pads with -1)."""
    X, Q = make(N, nq, d, dtype, seed=5)
    X[100:130] = X[99]            # a run of exact ties that straddles a band boundary for some queries
    idx = engine(X, dtype)
    res = idx.search_batch(Q, k)
    D, I = np.asarray(res.total_scores), np.asarray(res.total_indices)
    if dtype == "f32":
        Dr, Ir = c_oracle.flat_ip(Q, X, k)
        assert np.array_equal(I, Ir) and np.array_equal(D.view(np.uint32), Dr.view(np.uint32))
    else:
        Dr, Ir = flat_ip.flat_ip_search(Q, X, k)
        cmp = flat_ip.compare_results(Dr, Ir, D, I, rel_tol=REL_TOL_16, abs_tol=1e-4)
        assert cmp["real_mismatches"] == 0 and cmp["overlap"] >= MIN_OVERLAP and (I >= 0).sum() == (Ir >= 0).sum(), cmp
    valid = I >= 0
    assert all(len(set(row[v].tolist())) == v.sum() for row, v in zip(I, valid)), "an id was returned twice across bands"
    for row_d, v in zip(D, valid):
        n = int(v.sum())
        assert v[:n].all() and np.all(np.diff(row_d[:n]) <= 0), "valid hits form a descending prefix"


def test_large_k_cross_shard_merge():
    from viquae_b200.index import B200FlatIndex, merge_topk
    X, Q = make(6000, 12, 64, "bf16", seed=8)
    k = 400
    full = engine(X, "bf16")
    s_full, i_full = full.search_tensors(torch.from_numpy(Q).cuda(), k)
    parts = []
    for a, b in ((0, 2500), (2500, 6000)):
        sh = B200FlatIndex(d=64, dtype="bf16", id_offset=a)
        sh.add(X[a:b])
        parts.append(sh.search_tensors(torch.from_numpy(Q).cuda(), k))
    s_m, i_m = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    cmp = flat_ip.compare_results(s_full.cpu().numpy(), i_full.cpu().numpy(), s_m.cpu().numpy(), i_m.cpu().numpy(), rel_tol=1e-5)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] == 1.0, cmp


def test_lifecycle_no_leak_and_varying_batches():
    """Create / search / destroy many indexes and alternate batch sizes (workspace regrowth, plan changes):
    results stay correct and device memory returns to its starting level."""
    from viquae_b200.index import B200FlatIndex
    X, Q = make(20000, 300, 128, "bf16", seed=21)
    Dr, Ir = flat_ip.flat_ip_search(Q, X, 10)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for rep in range(12):
        idx = B200FlatIndex(d=128, dtype="bf16")
        idx.add(X)
        for nq in (1, 300, 17, 129, 64):
            res = idx.search_batch(Q[:nq], 10)
            cmp = flat_ip.compare_results(Dr[:nq], Ir[:nq], res.total_scores, res.total_indices, rel_tol=REL_TOL_16)
            assert cmp["real_mismatches"] == 0, (rep, nq, cmp)
        s, i = idx.search_tensors(torch.from_numpy(Q).cuda(), 10)
        assert np.array_equal(i.cpu().numpy()[:, 0], Ir[:, 0])
        idx.close()
        del idx, s, i
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 * 2 ** 20, f"device memory leaked: {(free0 - free1) / 2**20:.1f} MiB"


def test_cuda_graph_replay_equals_direct_search():
    """The whole multi-pass search is capture-safe (no host sync, no allocation): a CUDA-graph replay returns
    exactly what a direct call returns, for new queries written into the static input, and refuses a grown index."""
    from viquae_b200.index import B200FlatIndex
    X, Q = make(40000, 7, 128, "bf16", seed=31)
    idx = B200FlatIndex(d=128, dtype="bf16")
    idx.add(torch.from_numpy(X).cuda())
    cap = idx.capture_search(7, 50)
    for seed in (1, 2, 3):
        q = torch.from_numpy(make(8, 7, 128, "bf16", seed=seed)[1]).cuda()
        s0, i0 = idx.search_tensors(q, 50)
        s1, i1 = cap(q)
        torch.cuda.synchronize()
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
    idx.add(torch.from_numpy(X[:10]).cuda())
    with pytest.raises(RuntimeError, match="capture again"):
        cap.replay()


@pytest.mark.parametrize("dtype,nq", [("bf16", 3), ("bf16", 200), ("f32", 40), ("f32x3", 9)])
def test_massive_exact_ties_resolve_to_the_lowest_ids(dtype, nq):
    """Every row of the corpus identical (and a second block of identical, slightly better rows in the middle): all scores
    of a query tie exactly, thousands of tied keys straddle every cut — the primed thresholds (next float below the k-th
    score), the in-kernel prunes and the selection merges (second bitwise search over the row words) must still return the
    LOWEST ids, like the oracle's tie rule."""
    from viquae_b200.index import B200FlatIndex
    N, d, k = 70000, 64, 100
    X = np.full((N, d), 0.5, dtype=np.float32)
    X[30000:30040] = 0.75                      # 40 better rows (all tied among themselves), then 69,960 tied rows
    Q = np.abs(np.random.default_rng(3).standard_normal((nq, d)).astype(np.float32)) + 0.25
    if dtype == "bf16":
        Q = flat_ip.bf16_round(Q)
    idx = B200FlatIndex(d=d, dtype=dtype)
    idx.add(X)
    res = idx.search_batch(Q, k)
    I = np.asarray(res.total_indices)
    want = np.concatenate([np.arange(30000, 30040), np.arange(0, k - 40)])
    assert np.array_equal(I, np.tile(want, (nq, 1))), I[0][:50]
    D = np.asarray(res.total_scores)
    assert np.all(D[:, :40] > D[:, 40:41]) and np.all(D[:, 40:] == D[:, 40:41])
    idx.close()


def test_chained_small_batch_launches_equal_the_large_batch_path():
    """Batches of <= 128 queries chain their kernels with programmatic dependent launches (the next pass's contraction
    starts streaming while the previous pass's merge still runs; only its epilogue waits).  Many back-to-back calls of
    varying size, never synchronised in between, must return exactly what the CTA-pair path (plain launches) returns
    for the same queries: scores are the same TMEM accumulations and ties break by id on both paths."""
    from viquae_b200.index import B200FlatIndex
    X, Q = make(300000, 200, 128, "bf16", seed=41)
    idx = B200FlatIndex(d=128, dtype="bf16")
    idx.add(torch.from_numpy(X).cuda())
    Qd = torch.from_numpy(Q).cuda()
    s_ref, i_ref = idx.search_tensors(Qd, 100)          # 200 queries -> cta_group::2, no chaining
    outs = []
    rng = np.random.default_rng(0)
    for it in range(60):
        n = int(rng.choice([1, 2, 5, 8, 31, 32, 64, 128]))
        a = int(rng.integers(0, 200 - n + 1))
        outs.append((a, n, idx.search_tensors(Qd[a:a + n].contiguous(), 100)))
    torch.cuda.synchronize()
    for a, n, (s, i) in outs:
        assert torch.equal(i, i_ref[a:a + n]) and torch.equal(s, s_ref[a:a + n]), (a, n)
    # and through a CUDA graph, replayed back to back
    cap = idx.capture_search(5, 100, query_dtype=Qd.dtype)
    res = []
    for a in (0, 50, 100, 150, 195):
        s, i = cap(Qd[a:a + 5])
        res.append((a, s.clone(), i.clone()))
    torch.cuda.synchronize()
    for a, s, i in res:
        assert torch.equal(i, i_ref[a:a + 5]) and torch.equal(s, s_ref[a:a + 5]), a
    idx.close()


REL_TOL_F32X3 = 1e-5  # north star for the fp32 mode: scores within 1e-5 relative of the fp32 reference (measured: <= 3.3e-6, typically 6e-7)


@pytest.mark.parametrize("N,nq,d,k,dist,factory", [
    (5000, 70, 768, 100, "normal", "Flat"),
    (70001, 300, 768, 100, "normal", "Flat"),        # cta_group::2, several priming passes
    (40000, 130, 768, 100, "dpr", "Flat"),           # DPR-like: scores cluster near 71 +- 2, all products positive
    (30000, 40, 512, 100, "normal", "L2norm,Flat"),  # CLIP/ArcFace regime, scores in [-1, 1]
    (9000, 33, 100, 7, "normal", "Flat"),            # d not a multiple of 8: zero-padded segments
    (50, 3, 24, 100, "normal", "Flat"),              # N < k: -1 padding
    (30000, 20, 256, 300, "normal", "Flat"),         # k > 128: banded rounds
])
def test_f32x3_split_operand_mode_is_fp32_faithful(N, nq, d, k, dist, factory):
    """dtype="f32x3": fp32 embeddings as three bf16 terms, six tensor-core product terms, fp32 accumulation.
    Against the fp32 fixed-order oracle on the SAME fp32 inputs: scores within 1e-5 relative, ids identical
    except where two scores are closer than that (FAISS's own sgemm is no closer to the fixed-order sum)."""
    X, Q = make(N, nq, d, "f32", seed=17, dist=dist)
    X, Q = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Q, np.float32)  # the DPR-like family comes out as float64
    idx = engine(X, "f32x3", string_factory=factory)
    res = idx.search_batch(Q, k)
    D, I = np.asarray(res.total_scores), np.asarray(res.total_indices)
    if factory != "Flat":  # the oracle gets the engine's own normalised fp32 rows (same arithmetic as the f32 index)
        Xn = idx.reconstruct_n(0, None, "f32")
        ref = engine(X, "f32", string_factory=factory)
        assert np.array_equal(Xn.view(np.uint32), ref.reconstruct_n(0, None, "f32").view(np.uint32)), "split + join must be exact"
        Qn = Q / np.linalg.norm(Q, axis=1, keepdims=True)
        Dr, Ir = c_oracle.flat_ip(Qn.astype(np.float32), Xn, k)
        tol = REL_TOL_F32X3  # (the query normalisation also differs in the last bit between numpy and the device)
    else:
        assert np.array_equal(idx.reconstruct_n(0, None, "f32").view(np.uint32), X.view(np.uint32)), "split + join must be exact"
        Dr, Ir = c_oracle.flat_ip(Q, X, k)
        tol = REL_TOL_F32X3
    kk = min(k, N)
    floor = 1.0 if factory != "Flat" else float(np.sqrt(d))  # one standard deviation of a score: keeps near-zero scores meaningful
    # ids: the two top-k sets may differ only by rows whose score is within the tolerance of the k-th score (another
    # summation order swaps near-ties; FAISS's blocked sgemm would too)
    cmp = flat_ip.compare_results(Dr, Ir, D, I, rel_tol=tol, abs_tol=tol * floor)
    assert cmp["real_mismatches"] == 0 and cmp["overlap"] >= 0.995, cmp
    # scores, rank by rank (a swap of two near-equal neighbours leaves the sorted score lists within the tolerance)
    valid = Ir >= 0
    err = np.abs(D.astype(np.float64) - Dr) / np.maximum(np.abs(Dr), floor)
    assert float(err[valid].max()) <= tol, float(err[valid].max())
    # and every returned (id, score) pair is that row's true inner product
    Xs = idx.reconstruct_n(0, None, "f32").astype(np.float64)
    Qs = (Q / np.linalg.norm(Q, axis=1, keepdims=True) if factory != "Flat" else Q).astype(np.float64)
    got = I >= 0
    truth = np.einsum("qkd,qd->qk", Xs[np.where(got, I, 0)], Qs)
    err = np.abs(D - truth) / np.maximum(np.abs(truth), floor)
    assert float(err[got].max()) <= tol, float(err[got].max())
    assert np.array_equal(got, valid) and np.all(I[:, kk:] == -1) and np.all(np.diff(D[:, :kk], axis=1) <= 0)


def test_f32x3_save_load_and_device_queries(tmp_path):
    from viquae_b200.index import B200FlatIndex
    X, Q = make(3000, 9, 72, "f32", seed=23)
    idx = engine(X, "f32x3")
    path = str(tmp_path / "x3.index")
    idx.save(path)
    back = B200FlatIndex.load(path)
    assert back.dtype == "f32x3" and len(back) == 3000
    s0, i0 = idx.search_tensors(torch.from_numpy(Q).cuda(), 20)
    s1, i1 = back.search_tensors(torch.from_numpy(Q).cuda().to(torch.bfloat16).float(), 20)  # other query values: just runs
    s2, i2 = back.search_tensors(torch.from_numpy(Q).cuda(), 20)
    assert torch.equal(i0, i2) and torch.equal(s0, s2)
    with pytest.raises(ValueError):
        idx.reconstruct_n(0, None, "f32x3")
