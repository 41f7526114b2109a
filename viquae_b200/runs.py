"""Run / Qrels containers and the handful of metrics the search path reports.

The reference hands its results to `ranx` (``ranx.Run`` / ``ranx.Qrels`` / ``ranx.evaluate``,
meerqat/ir/search.py:485-505, fuse.py:42,236), which is not installable offline.  These light
containers keep the same observable contract for the hot path's callers:

* ``Run.run`` is ``dict[q_id -> dict[doc_id -> score]]`` whose inner dicts iterate in DESCENDING
  score order (consumers take the first key as top-1: ir/embedding.py:136, viz/html.py:48-51,
  train/data.py:795,935-936);
* ``save`` / ``from_file`` use the same JSON layout as ranx (a plain nested dict).

Metrics (``mrr@k``, ``precision@k``, ``hit_rate@k`` — the defaults at search.py:397) are evaluation
utilities outside the GPU path; they run on the host over the dicts.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional


def _sorted_desc(results: Dict[str, float]) -> Dict[str, float]:
    return dict(sorted(results.items(), key=lambda kv: -kv[1]))


class Run:
    def __init__(self, run: Optional[Dict[str, Dict[str, float]]] = None, name: Optional[str] = None):
        self.name = name
        self.run = {q: _sorted_desc(res) for q, res in (run or {}).items()}

    def keys(self):
        return self.run.keys()

    def __len__(self):
        return len(self.run)

    def __getitem__(self, q_id):
        return self.run[q_id]

    def to_dict(self):
        return self.run

    def save(self, path):
        with open(path, "wt") as f:
            json.dump(self.run, f)

    @classmethod
    def from_file(cls, path, name: Optional[str] = None) -> "Run":
        with open(path, "rt") as f:
            return cls(json.load(f), name=name)

    @classmethod
    def from_dict(cls, d, name: Optional[str] = None) -> "Run":
        return cls(d, name=name)


class Qrels:
    def __init__(self, qrels: Optional[Dict[str, Dict[str, int]]] = None, name: Optional[str] = None):
        self.name = name
        self.qrels = {q: dict(rels) for q, rels in (qrels or {}).items()}

    def keys(self):
        return self.qrels.keys()

    def __getitem__(self, q_id):
        return self.qrels[q_id]

    def to_dict(self):
        return self.qrels

    def save(self, path):
        with open(path, "wt") as f:
            json.dump(self.qrels, f)

    @classmethod
    def from_file(cls, path) -> "Qrels":
        with open(path, "rt") as f:
            return cls(json.load(f))

    @classmethod
    def from_dict(cls, d) -> "Qrels":
        return cls(d)


def _parse_metric(metric: str):
    name, _, cut = metric.partition("@")
    return name, (int(cut) if cut else None)


def evaluate(qrels, run, metrics, return_mean: bool = True):
    """Mean of mrr@k / precision@k / hit_rate@k / recall@k over the queries of ``qrels`` (queries missing from the
    run score 0, like ranx).  ``metrics`` may be a string or a list; returns a float or a dict."""
    q = qrels.qrels if isinstance(qrels, Qrels) else qrels
    r = run.run if isinstance(run, Run) else run
    single = isinstance(metrics, str)
    names = [metrics] if single else list(metrics)
    out = {}
    for metric in names:
        name, cut = _parse_metric(metric)
        vals = []
        for q_id, rels in q.items():
            relevant = {d for d, g in rels.items() if g > 0}
            ranked = list(r.get(q_id, {}).keys())
            if cut is not None:
                ranked = ranked[:cut]
            hits = [d in relevant for d in ranked]
            if name == "mrr":
                v = next((1.0 / (i + 1) for i, h in enumerate(hits) if h), 0.0)
            elif name == "precision":
                v = sum(hits) / float(cut if cut else max(len(ranked), 1))
            elif name == "hit_rate":
                v = 1.0 if any(hits) else 0.0
            elif name == "recall":
                v = sum(hits) / float(len(relevant)) if relevant else 0.0
            else:
                raise ValueError(f"unsupported metric {metric!r} (supported: mrr, precision, hit_rate, recall, each with @k)")
            vals.append(v)
        out[metric] = (sum(vals) / len(vals) if vals else 0.0) if return_mean else vals
    return out[names[0]] if single else out
