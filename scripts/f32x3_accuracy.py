"""Score accuracy of the fp32 modes against float64 dot products of the same fp32 inputs:
the fixed-order SIMT mode (dtype="f32", == oracle/flat_ip.c bit for bit) and the split-operand tensor-core
mode (dtype="f32x3").  python scripts/f32x3_accuracy.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from viquae_b200.index import B200FlatIndex

for dist in ("normal", "dpr"):
    rng = np.random.default_rng(3)
    N, nq, d, k = 200_000, 64, 768, 100
    if dist == "dpr":
        m = rng.standard_normal(d).astype(np.float32) * (8.0 / np.sqrt(d))
        X = m + 0.35 * rng.standard_normal((N, d), dtype=np.float32)
        Q = m * 8.9 + 0.35 * rng.standard_normal((nq, d), dtype=np.float32)
    else:
        X = rng.standard_normal((N, d), dtype=np.float32)
        Q = rng.standard_normal((nq, d), dtype=np.float32)
    out = {"dist": dist, "N": N, "nq": nq, "d": d, "k": k}
    ids = {}
    for dtype in ("f32", "f32x3", "bf16"):
        idx = B200FlatIndex(d=d, dtype=dtype)
        idx.add(X)
        r = idx.search_batch(Q, k)
        D, I = np.asarray(r.total_scores, np.float64), np.asarray(r.total_indices)
        truth = np.einsum("qkd,qd->qk", X[I].astype(np.float64), Q.astype(np.float64))
        rel = np.abs(D - truth) / np.abs(truth)
        out[dtype] = {"max_rel_err_vs_f64": float(rel.max()), "mean_rel_err_vs_f64": float(rel.mean())}
        ids[dtype] = I
        idx.close()
    out["ids_equal_f32_vs_f32x3"] = bool(np.array_equal(ids["f32"], ids["f32x3"]))
    out["top100_overlap_bf16_vs_f32"] = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(ids["f32"], ids["bf16"])]))
    print(json.dumps(out), flush=True)
