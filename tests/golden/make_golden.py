"""Generates the golden fixtures in tests/golden/ by running the REFERENCE's own Python code
(/root/reference/meerqat/ir/{search,fuse}.py, unmodified) in this container:

  fusion_golden.json  inputs (3 synthetic runs) -> outputs of the reference's default_minimum
                      (fuse.py:132-149) and of its numba gzmuv_norm (fuse.py:86-129)
  search_golden.npz / search_golden.json
                      a toy KB driven through the reference's KnowledgeBase / Searcher with the faiss-API
                      shim (tests/shims/faiss) backed by the CPU oracle: search_batch outputs, do_L2norm
                      inference, None-query handling, article->passage expansion, str doc ids, truncation at k

Run from the repo root:  python tests/golden/make_golden.py     (needs /root/reference)
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refstubs  # noqa: E402


def to_typed_run(run_dict):
    from numba import types
    from numba.typed import Dict
    outer = Dict.empty(key_type=types.unicode_type, value_type=types.DictType(types.unicode_type, types.float64))
    for q, res in run_dict.items():
        inner = Dict.empty(key_type=types.unicode_type, value_type=types.float64)
        for d, s in res.items():
            inner[d] = float(s)
        outer[q] = inner
    return outer


def make_fusion():
    fu = refstubs.load_reference_fuse()
    Run = sys.modules["ranx"].Run
    rng = np.random.default_rng(42)
    nq, k = 12, 10
    runs = []
    for r, (mu, sd, pool) in enumerate([(71.33, 2.167, 40), (0.3, 0.1, 25), (20.11, 5.85, 30)]):
        run = {}
        for q in range(nq):
            if r == 2 and q % 3 == 0:      # e.g. no face detected: the run has no result for this query
                run[f"q{q}"] = {}
                continue
            docs = rng.choice(pool, size=k, replace=False)
            scores = np.sort(rng.normal(mu, sd, size=k))[::-1]
            run[f"q{q}"] = {str(int(d)): float(s) for d, s in zip(docs, scores)}
        runs.append(run)
    # reference default_minimum mutates ranx.Run.run in place
    ref_runs = [Run() for _ in runs]
    for rr, run in zip(ref_runs, runs):
        rr.run = {q: dict(res) for q, res in run.items()}
        rr.name = "r"
    after_defmin = [r.run for r in fu.default_minimum(ref_runs)]
    # reference gzmuv_norm (numba) on the defmin-filled runs and on the raw runs
    gz_after_defmin, gz_raw = [], []
    for run in after_defmin:
        rr = Run(); rr.name = "r"; rr.run = to_typed_run(run)
        out = fu.gzmuv_norm(rr).run
        gz_after_defmin.append({q: dict(res) for q, res in out.items()})
    for run in runs:
        rr = Run(); rr.name = "r"; rr.run = to_typed_run(run)
        out = fu.gzmuv_norm(rr).run
        gz_raw.append({q: dict(res) for q, res in out.items()})
    with open(os.path.join(HERE, "fusion_golden.json"), "wt") as f:
        json.dump({"runs": runs, "after_defmin": after_defmin, "gzmuv_after_defmin": gz_after_defmin, "gzmuv_raw": gz_raw,
                   "weights_examples": [[0.3, 0.2, 0.2, 0.2], [0.5, 0.5], [0.6, 0.2, 0.2], [0.3, 0.1, 0.3, 0.3]]}, f)
    print("fusion_golden.json written")


def make_search():
    se = refstubs.load_reference_search()
    import datasets
    import faiss
    faiss.set_backend("oracle")
    rng = np.random.default_rng(7)
    n_art, d_txt, d_img, k = 400, 32, 16, 10
    txt = rng.standard_normal((n_art, d_txt)).astype(np.float32)
    img = (rng.standard_normal((n_art, d_img)) * 3).astype(np.float32)
    nq = 9
    q_txt = rng.standard_normal((nq, d_txt)).astype(np.float32)
    q_img = (rng.standard_normal((nq, d_img)) * 2).astype(np.float32)
    # article -> passages mapping (one2many), ~3 passages per article
    mapping, nxt = {}, 0
    for a in range(n_art):
        m = int(rng.integers(1, 5))
        mapping[str(a)] = list(range(nxt, nxt + m))
        nxt += m
    with tempfile.TemporaryDirectory() as tmp:
        kb_path = os.path.join(tmp, "kb")
        datasets.Dataset.from_dict({"text_emb": txt.tolist(), "image_emb": img.tolist()}).save_to_disk(kb_path)
        map_path = os.path.join(tmp, "article2passage.json")
        with open(map_path, "wt") as f:
            json.dump(mapping, f)
        kb_kwargs = {kb_path: {"index_mapping_path": map_path, "index_kwargs": {
            "dpr": {"column": "text_emb", "key": "q_text", "string_factory": "Flat", "metric_type": 0},
            "clip": {"column": "image_emb", "key": "q_image", "string_factory": "L2norm,Flat", "metric_type": 0}}}}
        qrels_path = os.path.join(tmp, "qrels.json")
        with open(qrels_path, "wt") as f:
            json.dump({}, f)
        searcher = se.Searcher(kb_kwargs, k=k, qrels=qrels_path, do_fusion=False)
        kb = searcher.kbs[kb_path]
        out = {"do_L2norm": {n: bool(ix.do_L2norm) for n, ix in kb.indexes.items()}}
        s1, i1 = kb.search_batch("dpr", q_txt, k=k)
        s2, i2 = kb.search_batch("clip", q_img, k=k)
        q_img_list = [None if j % 4 == 1 else q_img[j].tolist() for j in range(nq)]
        s3, i3 = kb.search_batch_if_not_None("clip", q_img_list, k=k)
        out["if_not_none_lengths"] = [len(x) for x in i3]
        batch = {"id": [f"q{j}" for j in range(nq)], "output": [{"original_answer": "", "answer": []}] * nq,
                 "q_text": q_txt.tolist(), "q_image": q_img_list}
        searcher(batch)
        out["runs"] = {name: {q: {doc: float(s) for doc, s in res.items()} for q, res in run.items()}
                       for name, run in searcher.runs.items()}
        out["mapping"] = mapping
        out["k"] = k
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), txt=txt, img=img, q_txt=q_txt, q_img=q_img,
                        dpr_scores=np.asarray(s1), dpr_ids=np.asarray(i1), clip_scores=np.asarray(s2), clip_ids=np.asarray(i2),
                        none_mask=np.array([q is None for q in q_img_list]))
    with open(os.path.join(HERE, "search_golden.json"), "wt") as f:
        json.dump(out, f)
    print("search_golden.{npz,json} written")


if __name__ == "__main__":
    assert refstubs.available(), "needs the reference checkout"
    make_fusion()
    make_search()
