"""Late fusion of the per-index runs — the ``meerqat.ir.fuse`` interface over the K5 CUDA kernels.

Same constructor / method names and kwargs as the reference's ``Fusion`` (meerqat/ir/fuse.py:162-237):
``Fusion(qrels, runs, norm="zmuv", method="wsum", output=None, defmin=False)``, ``.fit(metric)``,
``.test(best_params, metrics)``, plus the module-level ``default_minimum`` (fuse.py:132-149) and
``gzmuv_norm`` (fuse.py:86-129).  All score arithmetic (default-minimum fill, (g)zmuv / min-max
normalisation, weighted sum, descending sort) runs on the GPU in float64 through ``vq_fuse``; the
host only converts between the reference's dict runs and dense ``[n_runs, nq, k]`` arrays.  ``Fusion.fit``'s
grid search over weight vectors is one device pass too (``vq_fuse_grid``).

Supported: ``method="wsum"`` with ``norm in {"zmuv", "gzmuv", "min-max", None}`` — the combinations
the shipped experiment configs use.  ranx's other norms/methods raise ``NotImplementedError``.
"""
from __future__ import annotations

import itertools
import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from .runs import Qrels, Run, evaluate

NORM_CODES = {None: 0, "none": 0, "zmuv": 1, "gzmuv": 2, "min-max": 3}


def _as_dict(run) -> Dict[str, Dict[str, float]]:
    return run.run if hasattr(run, "run") else run


class DenseRuns:
    """Dense view of R runs over a common query list: scores f64 / ids i64 [R, nq, k_in], id -1 = absent."""

    def __init__(self, runs: Sequence):
        dicts = [_as_dict(r) for r in runs]
        self.q_ids: List = []
        seen = set()
        for d in dicts:
            for q in d:
                if q not in seen:
                    seen.add(q)
                    self.q_ids.append(q)
        all_docs = [doc for d in dicts for res in d.values() for doc in res]
        self.vocab = None
        try:
            as_int = {doc: int(doc) for doc in set(all_docs)}
            if any(v < 0 or v >= 2 ** 32 - 1 for v in as_int.values()) or len(set(as_int.values())) != len(as_int):
                raise ValueError
            self._encode = as_int
            self.str_ids = any(isinstance(doc, str) for doc in as_int)
        except (TypeError, ValueError):
            self.vocab = sorted(set(all_docs), key=str)
            self._encode = {doc: i for i, doc in enumerate(self.vocab)}
            self.str_ids = True
        R, nq = len(dicts), len(self.q_ids)
        k_in = max([len(res) for d in dicts for res in d.values()] + [1])
        self.scores = np.zeros((R, nq, k_in), dtype=np.float64)
        self.ids = np.full((R, nq, k_in), -1, dtype=np.int64)
        self.has_query = np.zeros((R, nq), dtype=bool)
        for r, d in enumerate(dicts):
            for qi, q in enumerate(self.q_ids):
                res = d.get(q)
                if res is None:
                    continue
                self.has_query[r, qi] = True
                for j, (doc, s) in enumerate(res.items()):
                    self.ids[r, qi, j] = self._encode[doc]
                    self.scores[r, qi, j] = s

    def decode(self, i: int):
        if self.vocab is not None:
            return self.vocab[i]
        return str(i) if self.str_ids else i


def fuse_dense(scores, ids, weights, norm: Optional[str] = "zmuv", defmin: bool = False, device: int = 0):
    """GPU fusion of dense runs.  scores/ids: array-likes or cuda tensors [R, nq, k_in]; returns cuda tensors
    (fused_scores f64 [nq, R*k_in] descending, fused_ids i64 [nq, R*k_in] (-1 padding), counts i32 [nq])."""
    import torch

    if norm not in NORM_CODES:
        raise NotImplementedError(f"norm {norm!r}: the B200 fusion kernel implements {sorted(k for k in NORM_CODES if k)} and None")
    dev = torch.device(f"cuda:{device}")
    s = torch.as_tensor(scores, dtype=torch.float64).to(dev).contiguous()
    i = torch.as_tensor(ids, dtype=torch.int64).to(dev).contiguous()
    R, nq, k_in = s.shape
    w = torch.as_tensor(np.asarray(weights, dtype=np.float64)).to(dev).contiguous()
    if w.numel() != R:
        raise ValueError(f"{R} runs but {w.numel()} weights")
    out_s = torch.empty((nq, R * k_in), dtype=torch.float64, device=dev)
    out_i = torch.empty((nq, R * k_in), dtype=torch.int64, device=dev)
    out_c = torch.zeros((nq,), dtype=torch.int32, device=dev)
    if nq == 0:
        return out_s, out_i, out_c
    lib = _lib.load()
    nbytes = lib.vq_fuse_workspace_bytes(R, nq, k_in)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.vq_fuse(s.data_ptr(), i.data_ptr(), w.data_ptr(), R, nq, k_in, int(bool(defmin)), NORM_CODES[norm],
                               out_s.data_ptr(), out_i.data_ptr(), out_c.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out_s, out_i, out_c


def _fuse_runs(runs: Sequence, weights, norm, defmin, name=None, device: int = 0) -> Run:
    dense = DenseRuns(runs)
    out_s, out_i, out_c = fuse_dense(dense.scores, dense.ids, weights, norm=norm, defmin=defmin, device=device)
    out_s, out_i, out_c = out_s.cpu().numpy(), out_i.cpu().numpy(), out_c.cpu().numpy()
    fused = Run(name=name)
    for qi, q in enumerate(dense.q_ids):
        n = int(out_c[qi])
        fused.run[q] = {dense.decode(int(d)): float(s) for d, s in zip(out_i[qi, :n], out_s[qi, :n])}  # already descending
    return fused


def default_minimum(runs, device: int = 0):
    """meerqat/ir/fuse.py:132-149 — per query, every run that has at least one result gets the documents it
    lacks (w.r.t. the union over runs) at its own per-query minimum.  Returns new runs (the reference mutates in
    place and returns the same list; callers use the return value, fuse.py:183-184)."""
    dicts = [_as_dict(r) for r in runs]
    out = []
    for r, run in enumerate(runs):
        one_hot = [1.0 if j == r else 0.0 for j in range(len(runs))]
        filled = _fuse_runs(dicts, one_hot, norm=None, defmin=True, name=getattr(run, "name", None), device=device)
        # the union pass yields every query of the union; keep this run's own queries, untouched where it was empty
        keep = Run(name=getattr(run, "name", None))
        for q, res in dicts[r].items():
            # one-hot weights: fused(d) = 1.0 * (own score | own minimum) + 0.0 * (other runs) — exact in float64.
            # A run without results for q is left empty (fuse.py:143-144), whatever the other runs hold.
            keep.run[q] = {} if len(res) == 0 else filled.run[q]
        out.append(keep)
    return out


def gzmuv_norm(run, device: int = 0) -> Run:
    """meerqat/ir/fuse.py:86-129 — (s - mu) / max(sigma, 1e-9) with mu, sigma over ALL scores of the run."""
    normed = _fuse_runs([run], [1.0], norm="gzmuv", defmin=False, name=getattr(run, "name", None), device=device)
    for q, res in _as_dict(run).items():  # keep queries without results
        normed.run.setdefault(q, {})
    return normed


NORMS = {"gzmuv": gzmuv_norm}


def weight_grid(n_runs: int, step: float = 0.1):
    """ranx.optimize_fusion's wsum search space: weights on a `step` grid that sum to 1."""
    ticks = int(round(1.0 / step))
    for combo in itertools.product(range(ticks + 1), repeat=n_runs):
        if sum(combo) == ticks:
            yield [c * step for c in combo]


class Fusion:
    """Optimise / apply late fusion (same surface as meerqat.ir.fuse.Fusion)."""

    def __init__(self, qrels: Union[str, Path, Qrels, List[str], None] = None, runs: Union[List[str], List[Run], None] = None,
                 norm: Union[Optional[str], List[Optional[str]]] = "zmuv",
                 method: Union[Optional[str], List[Optional[str]]] = "wsum",
                 output: Optional[str] = None, defmin: Optional[bool] = False, device: int = 0):
        if isinstance(qrels, Qrels) or qrels is None:
            self.qrels = qrels
        elif isinstance(qrels, (str, Path)):
            self.qrels = Qrels.from_file(qrels)
        elif isinstance(qrels, dict):
            self.qrels = Qrels(qrels)
        else:
            self.qrels = fuse_qrels(qrels)
        if not runs:
            raise ValueError("Fusion needs at least one run")
        self.runs = [r if isinstance(r, Run) else (Run(r) if isinstance(r, dict) else Run.from_file(r)) for r in runs]
        self.defmin = bool(defmin)
        self.device = device
        self.norm = norm
        self.method = method
        if output is not None:
            output = Path(output)
            output.mkdir(exist_ok=True)
        self.output = output

    @staticmethod
    def _check_method(method):
        if method not in ("wsum",):
            raise NotImplementedError(f"fusion method {method!r}: the B200 engine implements 'wsum' (the reference's default)")

    def _fused(self, weights, norm) -> Run:
        # defmin precedes normalisation (fuse.py:183-184 vs :199-200,224-225): the kernel applies them in that order
        return _fuse_runs(self.runs, weights, norm=norm, defmin=self.defmin, name="fusion", device=self.device)

    METRIC_CODES = {"mrr": 0, "precision": 1, "hit_rate": 2, "recall": 3}

    def fit(self, metric: str = "mrr@100"):
        """Finds the best weights on the 0.1 grid for every (norm, method) (fuse.py:193-217 -> ranx.optimize_fusion);
        writes ``{norm}_{method}_best_params.yaml``.  One ``vq_fuse_grid`` call per (norm, method): union +
        default-minimum + normalisation once, then all 11 / 66 / 286 weight vectors are scored on the device from the
        ranks of the relevant documents; the host only builds the dense arrays and reads back ``[n_weights]`` metric
        values and the index of the first best."""
        import torch

        if self.qrels is None:
            raise ValueError("Fusion.fit needs qrels")
        norms = [self.norm] if self.norm is None or isinstance(self.norm, str) else self.norm
        methods = [self.method] if self.method is None or isinstance(self.method, str) else self.method
        name, _, cut = metric.partition("@")
        if name not in self.METRIC_CODES:
            raise ValueError(f"unsupported metric {metric!r}: the device grid search scores {sorted(self.METRIC_CODES)} (@k)")
        cut = int(cut) if cut else 0
        dense = DenseRuns(self.runs)
        dev = torch.device(f"cuda:{self.device}")
        d_scores = torch.from_numpy(dense.scores).to(dev)
        d_ids = torch.from_numpy(dense.ids).to(dev)
        R, nq, k_in = dense.scores.shape
        qrels = self.qrels.qrels
        codes = [[dense._encode[d] for d, g in qrels.get(q, {}).items() if g > 0 and d in dense._encode] for q in dense.q_ids]
        max_rel = max([len(c) for c in codes] + [1])
        if max_rel > 256:
            raise NotImplementedError(f"a query has {max_rel} relevant documents among the retrieved ones; the device grid search holds 256 per query")
        rel = np.full((nq, max_rel), -1, dtype=np.int64)
        for qi, c in enumerate(codes):
            rel[qi, :len(c)] = c
        n_rel_total = np.array([sum(1 for g in qrels.get(q, {}).values() if g > 0) for q in dense.q_ids], dtype=np.int32)
        d_rel = torch.from_numpy(rel).to(dev)
        d_nrel = torch.from_numpy(n_rel_total).to(dev)
        grid = list(weight_grid(R))
        d_w = torch.tensor(grid, dtype=torch.float64, device=dev)
        W = len(grid)
        lib = _lib.load()
        ws = torch.empty(lib.vq_fuse_grid_workspace_bytes(R, nq, k_in, W) + 256, dtype=torch.uint8, device=dev)
        out_metric = torch.empty((W,), dtype=torch.float64, device=dev)
        out_best = torch.zeros((1,), dtype=torch.int32, device=dev)
        results = {}
        self.grid_reports = {}
        for norm in norms:
            if norm not in NORM_CODES:
                raise NotImplementedError(f"norm {norm!r}: the B200 fusion kernel implements {sorted(k for k in NORM_CODES if k)} and None")
            for method in methods:
                self._check_method(method)
                with torch.cuda.device(dev):
                    _lib.check(lib.vq_fuse_grid(d_scores.data_ptr(), d_ids.data_ptr(), R, nq, k_in, int(self.defmin), NORM_CODES[norm],
                                                d_w.data_ptr(), W, d_rel.data_ptr(), max_rel, d_nrel.data_ptr(), len(qrels),
                                                self.METRIC_CODES[name], cut, out_metric.data_ptr(), out_best.data_ptr(),
                                                ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
                scores = out_metric.cpu().numpy()
                best_i = int(out_best.item())
                best, best_score = grid[best_i], float(scores[best_i])
                self.grid_reports[(norm, method)] = list(zip(grid, scores.tolist()))
                best_params = {"weights": [round(w, 10) for w in best]}
                results[(norm, method)] = (best_params, best_score)
                print(f"Norm: {norm}, Method: {method}. Best parameters: {best_params}. {metric}={best_score:.4f}")
                if self.output is not None:
                    import yaml
                    with open(self.output / f"{norm}_{method}_best_params.yaml", "wt") as file:
                        yaml.dump(json.loads(json.dumps(best_params)), file)
        return results

    def test(self, best_params: dict, metrics: Optional[List[str]] = None) -> Run:
        """Applies ``best_params = {"weights": [...]}``; saves ``test_run.json``; returns the fused run."""
        if metrics is None:
            metrics = ["mrr@100", "precision@1", "precision@20", "hit_rate@20"]
        norm = self.norm if (self.norm is None or isinstance(self.norm, str)) else self.norm[0]
        method = self.method if (self.method is None or isinstance(self.method, str)) else self.method[0]
        self._check_method(method)
        combined_run = self._fused(best_params["weights"], norm)
        if self.output is not None:
            combined_run.save(self.output / "test_run.json")
        if metrics and self.qrels is not None:
            print({m: round(100 * v, 1) for m, v in evaluate(self.qrels, combined_run, metrics).items()})
        return combined_run


def fuse_qrels(qrels_paths) -> Qrels:
    """meerqat/ir/metrics.py:206-234 — union of several qrels files; contradictions raise."""
    final = {}
    for i, path in enumerate(qrels_paths):
        for q_id, rels in Qrels.from_file(path).qrels.items():
            tgt = final.setdefault(q_id, {})
            for doc_id, score in rels.items():
                if doc_id in tgt and tgt[doc_id] != score:
                    raise ValueError(f"{path} contradicts a prior Qrels (one of {list(qrels_paths[:i])}).\n"
                                     f"Got {score} and {tgt[doc_id]} for question '{q_id}' and document '{doc_id}'")
                tgt[doc_id] = score
    return Qrels(final)
