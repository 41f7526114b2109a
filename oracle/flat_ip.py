"""CPU oracle for the exact inner-product top-k search path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product (``viquae_b200``) never does and has no CPU fallback.

PARITY UNPINNED for the contraction itself: the arithmetic of this path lives in FAISS
(``faiss-gpu>=1.7.1``, /root/reference/requirements.txt:14, un-vendored, not installable here)
and the reference ships no golden vectors for it (its only test, tests/search.py:1-20, cannot
import; its golden runs are listed as stripped in .MISSING_LARGE_BLOBS:1-7).  What follows is a
restatement of FAISS ``IndexFlat(d, METRIC_INNER_PRODUCT)`` semantics as reached through

* meerqat/ir/search.py:43-46    ``L2norm``  (row-wise q / ||q||_2, no epsilon)
* meerqat/ir/search.py:135-146  ``KnowledgeBase.search_batch`` (float32 cast, optional L2norm)
* meerqat/ir/search.py:207-249  ``add_or_load_faiss_index`` (string_factory semantics)
* datasets/search.py:276-288, 369-385 (index construction; 2-D / C-contiguous checks;
  ``faiss_index.search(queries, k)``; ``indices.astype(int)``)

FAISS IndexFlatIP semantics restated (upstream ``knn_inner_product``):
  scores = Q . X^T in float32; per query the k largest, sorted descending; ids int64;
  a candidate replaces the heap top only if strictly greater than it, the heap starts at
  (-FLT_MAX, -1): so -inf / -FLT_MAX / NaN scores are never selected and rows with fewer than k
  selectable candidates are padded with (-FLT_MAX, -1).  FAISS leaves the order among exactly
  tied scores unspecified; this oracle (and the engine) fix it: lower id first.

The Python layer of the reference above FAISS *is* pinned: tests/golden/ holds outputs of the
reference's own unmodified ``KnowledgeBase``/``Searcher`` code driven through a faiss-API shim
backed by this oracle (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(np.finfo(np.float32).min)


def l2norm(x: np.ndarray) -> np.ndarray:
    """meerqat/ir/search.py:43-46 — zero rows become NaN exactly like the reference."""
    x = np.asarray(x, dtype=np.float32)
    norms = np.linalg.norm(x, axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        return x / norms


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (the 'same quantised inputs' rule)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    out = r.astype(np.uint32).view(np.float32).reshape(x.shape).copy()
    out[nan] = np.nan
    return out


def fp16_round(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def split3_bf16(x: np.ndarray):
    """The engine's VQ_F32X3 storage (viquae_b200/csrc/vq_prep.cu: split3): an fp32 value as three bf16 terms
    hi + mid + lo with every remainder exact in fp32, so the three terms sum to the value.  Not a reference
    function — FAISS holds plain fp32 (datasets/search.py:276-288); this restates OUR layout so its arithmetic can
    be checked on the CPU."""
    x = np.asarray(x, np.float32)
    hi = bf16_round(x)
    r1 = (x - hi).astype(np.float32)
    mid = bf16_round(r1)
    r2 = (r1 - mid).astype(np.float32)
    lo = bf16_round(r2)
    return hi, mid, lo


def f32x3_scores(Q, X) -> np.ndarray:
    """Scores of the six-term split contraction (hi*lo + lo*hi + mid*mid + hi*mid + mid*hi + hi*hi), every product
    exact and the sums in float64: isolates the error of the three dropped terms from accumulation rounding."""
    qh, qm, ql = (t.astype(np.float64) for t in split3_bf16(Q))
    xh, xm, xl = (t.astype(np.float64) for t in split3_bf16(X))
    return qh @ xl.T + ql @ xh.T + qm @ xm.T + qh @ xm.T + qm @ xh.T + qh @ xh.T


def _select_topk(scores: np.ndarray, ids: np.ndarray, k: int):
    """Exact top-k of one row: descending score, ties -> lower id; unselectable values dropped.

    scores float32/float64 [m], ids int64 [m].  Returns (<=k scores, <=k ids).
    """
    ok = scores > NEG_FLT_MAX  # drops NaN, -inf and -FLT_MAX (strict '>' of the FAISS heap)
    if not ok.all():
        scores, ids = scores[ok], ids[ok]
    m = scores.shape[0]
    if m > k:
        # threshold = k-th largest value; keep everything above it plus lowest-id ties
        part = np.partition(scores, m - k)
        t = part[m - k]
        above = scores > t
        n_above = int(above.sum())
        eq = np.flatnonzero(scores == t)
        eq = eq[np.argsort(ids[eq], kind="stable")][: k - n_above]
        keep = np.concatenate([np.flatnonzero(above), eq])
        scores, ids = scores[keep], ids[keep]
    order = np.lexsort((ids, -scores.astype(np.float64)))
    return scores[order], ids[order]


def flat_ip_search(Q, X, k: int, *, l2norm_queries: bool = False, id_offset: int = 0,
                   block_q: int = 4096, block_x: int = 65536, dtype=np.float32):
    """Exact IP top-k, blocked like FAISS's sgemm path (4096-query x corpus-row blocks).

    Q [nq,d], X [N,d] -> (D float32 [nq,k] descending, I int64 [nq,k]); padding (-FLT_MAX, -1).
    ``dtype=np.float64`` gives the referee used to classify near-ties.
    """
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    if Q.ndim != 2:
        raise ValueError("Shape of query must be 2D")  # datasets/search.py:380-381
    if X.ndim != 2 or (X.shape[0] and X.shape[1] != Q.shape[1]):
        raise ValueError("dimension mismatch")
    if l2norm_queries:
        Q = l2norm(Q)
    nq, N = Q.shape[0], X.shape[0]
    D = np.full((nq, k), NEG_FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    Qc, Xc = Q.astype(dtype, copy=False), X.astype(dtype, copy=False)
    for q0 in range(0, nq, block_q):
        q1 = min(nq, q0 + block_q)
        best_s = [np.empty(0, dtype) for _ in range(q1 - q0)]
        best_i = [np.empty(0, np.int64) for _ in range(q1 - q0)]
        for x0 in range(0, N, block_x):
            x1 = min(N, x0 + block_x)
            S = Qc[q0:q1] @ Xc[x0:x1].T
            ids = np.arange(x0, x1, dtype=np.int64)
            for r in range(q1 - q0):
                s, i = _select_topk(S[r], ids, k)
                if best_s[r].size:
                    s, i = _select_topk(np.concatenate([best_s[r], s]), np.concatenate([best_i[r], i]), k)
                best_s[r], best_i[r] = s, i
        for r in range(q1 - q0):
            n = best_s[r].size
            D[q0 + r, :n] = best_s[r].astype(np.float32)
            I[q0 + r, :n] = best_i[r] + id_offset
    return D, I


def flat_ip_search_fast(Q, X, k: int, *, block_q: int = 4096, block_x: int = 262144, threads=None):
    """Throughput-oriented CPU baseline of the same path (used by bench.py's cpu_baseline and
    ``--impl reference`` legs): torch/MKL sgemm on all host cores + torch.topk per block + running
    merge, i.e. FAISS's blocked-sgemm + k-selection structure.  Results equal ``flat_ip_search``
    up to the order of exactly tied scores."""
    import torch

    if threads:
        torch.set_num_threads(int(threads))
    Qt = torch.as_tensor(np.ascontiguousarray(Q, dtype=np.float32))
    Xt = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32))
    nq, N = Qt.shape[0], Xt.shape[0]
    D = torch.full((nq, k), float(NEG_FLT_MAX))
    I = torch.full((nq, k), -1, dtype=torch.int64)
    for q0 in range(0, nq, block_q):
        q1 = min(nq, q0 + block_q)
        bs = torch.full((q1 - q0, k), float(NEG_FLT_MAX))
        bi = torch.full((q1 - q0, k), -1, dtype=torch.int64)
        for x0 in range(0, N, block_x):
            x1 = min(N, x0 + block_x)
            S = Qt[q0:q1] @ Xt[x0:x1].float().T
            kk = min(k, x1 - x0)
            s, i = torch.topk(S, kk, dim=1)
            cs = torch.cat([bs, s], dim=1)
            ci = torch.cat([bi, i + x0], dim=1)
            s2, o = torch.topk(cs, k, dim=1)
            bs, bi = s2, torch.gather(ci, 1, o)
        D[q0:q1], I[q0:q1] = bs, bi
    return D.numpy(), I.numpy()


def merge_topk(scores_list, ids_list, k: int):
    """k-way merge of per-shard top-k lists: top-k of the union, same tie rule."""
    S = np.concatenate([np.asarray(s, np.float32) for s in scores_list], axis=1)
    I = np.concatenate([np.asarray(i, np.int64) for i in ids_list], axis=1)
    nq = S.shape[0]
    D = np.full((nq, k), NEG_FLT_MAX, np.float32)
    O = np.full((nq, k), -1, np.int64)
    for r in range(nq):
        valid = I[r] >= 0
        s, i = _select_topk(S[r][valid], I[r][valid], k)
        D[r, : s.size], O[r, : s.size] = s, i
    return D, O


def compare_results(D_ref, I_ref, D_got, I_got, *, rel_tol: float, abs_tol: float = 0.0):
    """Classify differences between an engine result and the oracle (tie-aware).

    Returns dict(max_rel_err, id_mismatch_rows, real_mismatches, overlap) where a mismatch is
    'real' only if the sets of ids differ by an element whose oracle score is farther than the
    tolerance from the k-th oracle score (i.e. not a tie / near-tie at the boundary) or if
    position-wise scores differ by more than the tolerance.
    """
    D_ref, D_got = np.asarray(D_ref, np.float64), np.asarray(D_got, np.float64)
    I_ref, I_got = np.asarray(I_ref), np.asarray(I_got)
    denom = np.maximum(np.abs(D_ref), 1e-30)
    valid = I_ref >= 0
    rel = np.where(valid, np.abs(D_ref - D_got) / denom, 0.0)
    max_rel = float(rel.max()) if rel.size else 0.0
    nq, k = I_ref.shape
    real = 0
    overlap = 0
    mism_rows = 0
    for r in range(nq):
        a, b = set(I_ref[r][I_ref[r] >= 0].tolist()), set(I_got[r][I_got[r] >= 0].tolist())
        overlap += len(a & b)
        if not np.array_equal(I_ref[r], I_got[r]):
            mism_rows += 1
        if a != b:
            kth = D_ref[r][valid[r]].min() if valid[r].any() else 0.0
            tol = max(abs_tol, rel_tol * abs(kth))
            ref_map = dict(zip(I_ref[r].tolist(), D_ref[r].tolist()))
            got_map = dict(zip(I_got[r].tolist(), D_got[r].tolist()))
            for i in a - b:
                if ref_map[i] - kth > tol:
                    real += 1
            for i in b - a:
                if kth - got_map[i] > tol:
                    real += 1
    total = int(valid.sum())
    return dict(max_rel_err=max_rel, id_mismatch_rows=mism_rows, real_mismatches=real,
                overlap=(overlap / total) if total else 1.0)
