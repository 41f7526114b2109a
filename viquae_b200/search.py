"""``meerqat.ir.search``'s index abstraction and Searcher over the B200 engine.

Mirrors, for the dense exact-IP path only, the reference's public surface
(meerqat/ir/search.py): ``L2norm`` (:43-46), ``IndexKind`` (:49-52) with the new member ``B200``,
``Index`` (:55-75), ``KnowledgeBase`` (:81-293: ``search_batch`` :135-146,
``search_batch_if_not_None`` :148-171, ``add_or_load_index`` :173-205), ``Searcher`` (:296-459)
and ``dataset_search`` (:462-524).  An experiment config selects the engine with
``"kind": "B200"`` in ``index_kwargs`` (enum-name lookup like search.py:192); ``kind`` absent or
``"FAISS"`` with ``string_factory`` in {None + metric_type 0, "Flat", "L2norm,Flat"} is routed to
the same engine (that is the configuration it replaces).  ES / PYSERINI kinds are the
reference's sparse retrieval and are out of scope: selecting them raises.

Out of scope here and therefore injectable: relevance extension by answer-string matching
(``find_relevant``, meerqat/ir/metrics.py:79-124) — pass ``find_relevant=`` to ``Searcher``.
"""
from __future__ import annotations

import enum
import json
import warnings
from pathlib import Path
from typing import Optional

import numpy as np

from . import _lib
from .fuse import Fusion
from .index import B200FlatIndex, METRIC_INNER_PRODUCT, parse_string_factory
from .postprocess import CSRMapping, expand_hits
from .runs import Qrels, Run, evaluate

LEGACY_KEYS = ("es", "kind_str", "normalization", "interpolation_weight")  # older config dialect


def L2norm(queries, device: int = 0):
    """Row-wise ``q / ||q||_2`` (meerqat/ir/search.py:43-46) on the GPU; a zero row becomes NaN like numpy's."""
    import torch

    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim != 2:
        raise ValueError("Expects a batch of vectors of the same dimension")
    dev = torch.device(f"cuda:{device}")
    src = torch.from_numpy(q).to(dev)
    dst = torch.empty_like(src)
    if q.shape[0]:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().vq_prep_rows(src.data_ptr(), _lib.F32, dst.data_ptr(), _lib.F32, q.shape[0], q.shape[1], 1,
                                                torch.cuda.current_stream(dev).cuda_stream))
    return dst.cpu().numpy()


class IndexKind(enum.Enum):
    FAISS = 0
    ES = 1
    PYSERINI = 2
    B200 = 3


class Index:
    """Record about one index of a KB (meerqat/ir/search.py:55-75): the dataset column holding the queries
    (``key``), the ``kind`` and whether queries get L2-normalised."""

    def __init__(self, key, kind=IndexKind.B200, do_L2norm=False):
        self.key = key
        self.kind = kind
        self.do_L2norm = do_L2norm


def _dist_world():
    """(world_size, rank) of the default process group, (1, 0) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except Exception:
        pass
    return 1, 0


def _load_from_disk(path):
    from datasets import load_from_disk
    return load_from_disk(path)


class KnowledgeBase:
    """A KB indexed by one or several B200 indexes.  Same parameters as the reference (search.py:81-122);
    additionally ``dataset=`` accepts an in-memory HF Dataset / dict of column -> array instead of ``kb_path``."""

    _warned_default_dtype = False

    def __init__(self, kb_path=None, index_mapping_path=None, many2one=None, index_kwargs={}, es_client=None, load_dataset=True,
                 dataset=None):
        if dataset is not None:
            self.dataset = dataset
        elif load_dataset:
            self.dataset = _load_from_disk(kb_path)
        else:
            self.dataset = None
        self.es_client = es_client
        self.indexes = {}
        self._engines = {}
        if index_mapping_path is None:
            self.index_mapping = None
        else:
            with open(index_mapping_path, "rt") as file:
                self.index_mapping = {int(k): v for k, v in json.load(file).items()}  # JSON has no integer keys
        if many2one not in (None, "max"):
            raise ValueError(f"Invalid value for many2one: '{many2one}'. Choose from {{None, 'max'}}")
        self.many2one = many2one
        for index_name, index_kwarg in index_kwargs.items():
            self.add_or_load_index(index_name=index_name, **index_kwarg)

    # ------------------------------------------------------------------ index construction
    def add_or_load_index(self, column=None, index_name=None, kind=None, key=None, **index_kwarg):
        """Dispatch on ``kind`` (enum member NAME, search.py:189-192).  ``column=None`` only records the index."""
        kind = IndexKind.B200 if kind is None else IndexKind[kind]
        if index_name is None:
            index_name = column
        if kind in (IndexKind.ES, IndexKind.PYSERINI):
            raise NotImplementedError(f"index kind {kind.name} (sparse retrieval) is outside the scope of viquae_b200; "
                                      "use the reference for BM25 and this engine for the dense indexes")
        do_L2norm = self.add_or_load_b200_index(column, index_name=index_name, **index_kwarg)
        self.indexes[index_name] = Index(key=key, kind=IndexKind.B200, do_L2norm=do_L2norm)

    # the reference's FAISS handler name, kept so that callers of add_or_load_faiss_index keep working
    def add_or_load_faiss_index(self, column, index_name=None, **kwargs):
        return self.add_or_load_b200_index(column, index_name=index_name, **kwargs)

    def _column_rows(self, column, start, stop):
        ds = self.dataset
        if hasattr(ds, "with_format"):  # HF Dataset: numpy view of one column (arrow_dataset.py:6259 does the same)
            return np.asarray(ds.with_format("numpy", columns=[column])[start:stop][column], dtype=np.float32)
        return np.asarray(ds[column][start:stop], dtype=np.float32)

    def add_or_load_b200_index(self, column, index_name=None, load=False, save_path=None, string_factory=None, device=None,
                               metric_type=None, dtype=None, batch_size=65536, file=None, variant=None, **kwargs):
        """Builds (or loads) a ``B200FlatIndex`` over ``dataset[column]`` and registers it.

        ``string_factory`` / ``metric_type`` / ``device`` / ``load`` / ``save_path`` mean what they mean in
        ``add_or_load_faiss_index`` (search.py:207-249); ``dtype`` picks the storage type ("bf16" | "f16" | "f32" |
        "f32x3").  ``device``: ``None`` or an int -> one GPU; ``-1`` (all GPUs) or a list of GPUs -> the corpus is
        ROW-SHARDED over one process per GPU (run under ``torchrun``; the reference would replicate it,
        datasets/search.py:329-340) and every search returns the merged result on every rank.
        Returns ``do_L2norm`` inferred from ``string_factory``."""
        for legacy in LEGACY_KEYS:
            if legacy in kwargs:
                v = kwargs.pop(legacy)
                if legacy == "es" and v:
                    raise NotImplementedError("legacy key es=true selects ElasticSearch, which is out of scope")
                warnings.warn(f"ignoring legacy index option {legacy}={v!r}")
        if kwargs:
            raise TypeError(f"add_or_load_b200_index() got unexpected keyword arguments {sorted(kwargs)}")
        if string_factory is None and metric_type is None:
            metric_type = METRIC_INNER_PRODUCT if column is None else metric_type
        do_L2norm = parse_string_factory(string_factory, metric_type) if (column is not None or string_factory) else False
        if column is None:
            return do_L2norm
        if dtype is None:
            # A config that only swaps `kind` must get the runs FAISS would have produced: fp32 embeddings, fp32-faithful
            # scores.  "f32x3" does that on the tensor cores (scores within 1e-5 relative of an fp32 sgemm, 12 bytes per
            # element on the device); the 16-bit storage types are a choice the config has to spell out.
            dtype = "f32x3"
            if not KnowledgeBase._warned_default_dtype:
                KnowledgeBase._warned_default_dtype = True
                warnings.warn('viquae_b200: no "dtype" in index_kwargs -> "f32x3" (fp32-faithful, what the FAISS index it replaces computes). '
                              '"dtype": "bf16" stores 2 bytes per element and searches ~6x faster; its scores are those of the '
                              'bf16-rounded embeddings (within 2e-3 relative, ~0.5 % of a top-100 changes at near-ties).')
        from .sharded import ShardedB200Index, resolve_devices, shard_range
        world, rank = _dist_world()
        dev, sharded = resolve_devices(device, world, rank, _lib.load().vq_device_count())
        n = len(self.dataset) if self.dataset is not None else None
        r0, r1 = shard_range(n, rank, world) if (sharded and n is not None) else (0, n)
        if load:
            path = file or save_path
            if path is None:
                raise ValueError("load=True needs `file` (or `save_path`) pointing at a saved index")
            if sharded:
                path = f"{path}.shard{rank}of{world}"  # what ShardedB200Index.save wrote
            engine = B200FlatIndex.load(path, device=dev, variant=variant, dtype=dtype)
            # The file only knows what its writer knew.  With `device` set the reference strips "L2norm" from the
            # factory, pre-normalises the column and saves a plain Flat index (search.py:237-248), and normalises the
            # queries in Python because the CONFIG says "L2norm,Flat" (Index.do_L2norm, search.py:144-145): the
            # config's flag wins whenever it asks for normalisation.
            engine.do_L2norm = bool(engine.do_L2norm or do_L2norm)
            if sharded:
                engine = ShardedB200Index(engine)
            if self.dataset is not None and len(engine) != len(self.dataset):
                raise ValueError(f"Index size should match Dataset size, but Index '{index_name}' at {path} has "
                                 f"{len(engine)} elements while the dataset has {len(self.dataset)} examples.")
        else:
            engine = B200FlatIndex(dtype=dtype, device=dev, string_factory=string_factory or "Flat", capacity=r1 - r0, variant=variant,
                                   id_offset=r0)
            if hasattr(self.dataset, "with_format"):  # HF Dataset: zero-copy Arrow chunks where possible
                from .ingest import arrow_column_blocks
                for block in arrow_column_blocks(self.dataset, column, max_rows=batch_size, start=r0, stop=r1):
                    engine.add(block)
            else:
                for i in range(r0, r1, batch_size):
                    engine.add(self._column_rows(column, i, min(r1, i + batch_size)))
            if len(engine) != r1 - r0:
                raise RuntimeError(f"ingested {len(engine)} rows of column '{column}' but rows [{r0}, {r1}) were expected")
            if sharded:
                engine = ShardedB200Index(engine)  # rank r holds rows [r0, r1); len(engine) is the whole corpus
            if save_path is not None:
                engine.save(save_path)
        self._engines[index_name] = engine
        if hasattr(self.dataset, "_indexes"):  # Dataset.search_batch(index_name, ...) dispatches to the engine
            self.dataset._indexes[index_name] = engine
        return do_L2norm

    def csr_mapping(self, device: int = 0):
        """The article -> passages mapping as device CSR arrays (built once per device); None without a mapping."""
        if self.index_mapping is None:
            return None
        cache = self.__dict__.setdefault("_csr", {})
        if device not in cache:
            n_rows = max([max(self.index_mapping) + 1 if self.index_mapping else 0] + [len(e) for e in self._engines.values()])
            cache[device] = CSRMapping(self.index_mapping, n_rows=n_rows, device=device)
        return cache[device]

    # ------------------------------------------------------------------ search
    def search_batch(self, index_name, queries, k=100):
        """``(scores, indices)`` of the k best KB rows per query (search.py:135-146): float32 [nq, k] descending,
        int [nq, k]; the float32 cast and the optional L2 normalisation of the queries happen on the device."""
        if index_name not in self.indexes:
            from .index import BaseIndex  # noqa: F401
            raise KeyError(f"Index with index_name '{index_name}' not initialized yet. Please make sure that you call "
                           "`add_or_load_index` first.")
        engine = self._engines[index_name]
        queries = np.array(queries, dtype=np.float32)
        res = engine.search_batch(queries, k=k)  # the engine normalises when the index was built with L2norm
        return res.total_scores, res.total_indices

    def search_batch_if_not_None(self, index_name, queries, k=100):
        """Filters out ``None`` queries (e.g. no face detected), searches the rest, scatters the results back;
        ``None`` rows yield ``[]`` (search.py:148-171)."""
        scores_batch = [[] for _ in queries]
        indices_batch = [[] for _ in queries]
        kept = [i for i, query in enumerate(queries) if query is not None]
        if not kept:
            return scores_batch, indices_batch
        scores, indices = self.search_batch(index_name, [queries[i] for i in kept], k=k)
        for j, i in enumerate(kept):
            scores_batch[i] = scores[j]
            indices_batch[i] = indices[j]
        return scores_batch, indices_batch


class Searcher:
    """Aggregates KnowledgeBases, searches every index for each batch and accumulates runs / qrels
    (meerqat/ir/search.py:296-459)."""

    def __init__(self, kb_kwargs, k=100, reference_kb_path=None, reference_key="passage", qrels=None, request_timeout=1000,
                 es_client_kwargs={}, fusion_kwargs={}, metrics_kwargs={}, do_fusion=None, qnonrels=None, find_relevant=None,
                 gpu_postprocess=True):
        self.k = k
        self.gpu_postprocess = gpu_postprocess
        self.kbs = {}
        self.qrels = {}
        if qrels is not None:
            with open(qrels, "rt") as file:
                self.qrels = json.load(file)
        self.qnonrels = {}
        if qnonrels is not None:
            with open(qnonrels, "rt") as file:
                self.qnonrels = json.load(file)
        self.runs = {}
        resolved = {}
        for kb_path, kb_kwarg in kb_kwargs.items():
            rp = Path(kb_path).expanduser().resolve()
            if rp in resolved:
                raise ValueError(f"'{kb_path}' and '{resolved[rp]}' resolve to the same path")
            resolved[rp] = kb_path
            kb = KnowledgeBase(kb_path, **kb_kwarg)
            self.kbs[kb_path] = kb
            index_names = kb.indexes.keys()
            assert not (index_names & self.runs.keys()), "All KBs should have unique index names"
            for index_name in index_names:
                self.runs[index_name] = {}
        assert not ({"search", "fusion"} & self.runs.keys()), "'search', 'fusion' are reserved names"
        self.do_fusion = True if (do_fusion is None and len(self.runs) > 1) else do_fusion
        if self.do_fusion:
            assert len(self.runs) > 1
        self.find_relevant = find_relevant
        if reference_kb_path is None:
            assert qrels is not None, "At least one of {reference_kb_path, qrels} should be provided"
            warnings.warn("Didn't get a reference KB -> will not be able to extend the annotation coverage "
                          "so results should be interpreted carefully.\n")
            self.reference_kb = None
        else:
            if find_relevant is None:
                raise NotImplementedError("reference_kb_path needs answer-string relevance matching (meerqat.ir.metrics."
                                          "find_relevant), which is outside this engine: pass find_relevant=<callable>")
            kb = _load_from_disk(reference_kb_path)
            self.reference_kb = kb.remove_columns([c for c in kb.column_names if c != reference_key])
        self.reference_key = reference_key
        self.fusion_kwargs = fusion_kwargs
        self.metrics_kwargs = dict(metrics=["mrr@100", "precision@1", "precision@20", "hit_rate@20"])
        self.metrics_kwargs.update(metrics_kwargs)

    def _record(self, kb, run_q, scores, indices):
        """Post-processing of one query's hits (search.py:413-440): article -> passage expansion, many2one,
        str doc ids, stop once k documents are recorded."""
        for score, i in zip(scores, indices):
            i = int(i)
            if kb.index_mapping is not None:
                penalty = 0.0
                for j in kb.index_mapping[i]:
                    j = str(j)
                    if kb.many2one is None:
                        run_q[j] = score - penalty  # one2many: overwrite; passages of an article keep mapping order
                        penalty += 1e-8
                    elif j not in run_q or run_q[j] < score:  # many2one == 'max'
                        run_q[j] = score
            else:
                run_q[str(i)] = score
            if len(run_q) >= self.k:
                break

    def _search_index(self, kb, index_name, index, batch):
        """One index, one batch -> fills self.runs[index_name].  Device path: search -> K6 expansion -> one D2H
        copy; `None` queries (search.py:148-171) simply get an empty run."""
        import torch

        queries = batch[index.key]
        run = self.runs[index_name]
        kept = [i for i, query in enumerate(queries) if query is not None]
        for q_id in batch["id"]:
            run.setdefault(q_id, {})
        if not kept:
            return
        engine = kb._engines[index_name]
        fresh = all(len(run[batch["id"][i]]) == 0 for i in kept)
        if not (getattr(self, "gpu_postprocess", True) and fresh):
            scores, indices = kb.search_batch(index_name, [queries[i] for i in kept], k=self.k)
            for j, i in enumerate(kept):
                self._record(kb, run[batch["id"][i]], scores[j], indices[j])
            return
        q = np.array([queries[i] for i in kept], dtype=np.float32)
        if q.ndim != 2:
            raise ValueError("Shape of query must be 2D")
        dev = torch.device(f"cuda:{engine.device}")
        s, ids = engine.search_tensors(torch.from_numpy(q).to(dev), self.k)
        try:
            out_s, out_i, out_c = expand_hits(s, ids, kb.csr_mapping(engine.device), kb.many2one, self.k)
        except _lib.EngineError as e:
            if e.code != _lib.ERR_UNSUPPORTED:
                raise
            # an article with more passages than the kernel's shared-memory run can hold: the reference's loop on the host
            scores, indices = s.cpu().numpy(), ids.cpu().numpy()
            for j, i in enumerate(kept):
                self._record(kb, run[batch["id"][i]], scores[j], indices[j])
            return
        out_s, out_i, out_c = out_s.cpu().numpy(), out_i.cpu().numpy(), out_c.cpu().numpy()
        for j, i in enumerate(kept):
            n = int(out_c[j])
            run[batch["id"][i]].update({str(int(d)): float(v) for d, v in zip(out_i[j, :n], out_s[j, :n])})

    def __call__(self, batch):
        """Search a batch with every index of every KB; ``batch`` is an HF batch dict with ``id``, the query
        columns named by each index's ``key`` and (only with ``find_relevant``) ``output``."""
        n = len(batch["id"])
        outputs = batch.get("output", [None] * n)
        for kb in self.kbs.values():
            for index_name, index in kb.indexes.items():
                self._search_index(kb, index_name, index, batch)
                if self.reference_kb is None:
                    continue
                for q_id, gt in zip(batch["id"], outputs):
                    run_q = self.runs[index_name][q_id]
                    self.qrels.setdefault(q_id, {})
                    self.qnonrels.setdefault(q_id, {})
                    retrieved = run_q.keys() - (self.qrels[q_id].keys() | self.qnonrels[q_id].keys())
                    _, relevant = self.find_relevant(retrieved, gt["original_answer"], gt["answer"], self.reference_kb,
                                                     reference_key=self.reference_key)
                    self.qrels[q_id].update({str(i): 1 for i in relevant})
                    self.qnonrels[q_id].update({i: 0 for i in retrieved - self.qrels[q_id].keys()})
        return batch


def _iter_batches(dataset, batch_size):
    n = len(dataset)
    for i in range(0, n, batch_size):
        yield dataset[i:i + batch_size]


def dataset_search(dataset, k=100, metric_save_path=None, map_kwargs={}, **kwargs):
    """Searcher over a whole dataset, then metrics and (for several indexes) fusion (search.py:462-524).
    Returns the searcher (the reference returns None; everything else is written to ``metric_save_path``)."""
    searcher = Searcher(k=k, **kwargs)
    if hasattr(dataset, "map"):
        dataset.map(searcher, batched=True, **map_kwargs)
    else:
        for batch in _iter_batches(dataset, map_kwargs.get("batch_size", 256)):
            searcher(batch)
    searcher.qrels = Qrels(searcher.qrels)
    for name, run in searcher.runs.items():
        searcher.runs[name] = Run({q: {d: float(s) for d, s in res.items()} for q, res in run.items()}, name=name)
    # under torchrun (row-sharded indexes) every rank holds the same runs: rank 0 alone writes the files
    if _dist_world()[1] != 0:
        metric_save_path = None
    if metric_save_path is not None:
        metric_save_path = Path(metric_save_path)
        metric_save_path.mkdir(exist_ok=True)
        searcher.qrels.save(metric_save_path / "qrels.json")
        if searcher.qnonrels is not None:
            with open(metric_save_path / "qnonrels.json", "wt") as file:
                json.dump(searcher.qnonrels, file)
        for index_name, run in searcher.runs.items():
            run.save(metric_save_path / f"{index_name}.json")
    report = {name: evaluate(searcher.qrels, run, searcher.metrics_kwargs["metrics"]) for name, run in searcher.runs.items()}
    if _dist_world()[1] == 0:
        print(report)
    if metric_save_path is not None:
        with open(metric_save_path / "metrics.json", "wt") as file:
            json.dump(report, file)
    if searcher.do_fusion:
        fusion_kwargs = dict(searcher.fusion_kwargs)  # the reference pops from the Searcher's dict (quirk 7); we copy
        subcommand = fusion_kwargs.pop("subcommand")
        subcommand_kwargs = fusion_kwargs.pop("subcommand_kwargs", {})
        fuser = Fusion(qrels=searcher.qrels, runs=list(searcher.runs.values()), output=metric_save_path, **fusion_kwargs)
        searcher.fusion_result = getattr(fuser, subcommand)(**subcommand_kwargs)
    return searcher


def main(argv=None):
    """``python -m viquae_b200.search <dataset> <config> [--k=<k>] [--metrics=<path>]`` (search.py:1-15, 527-540)."""
    import argparse

    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("dataset")
    ap.add_argument("config")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--metrics", default=None)
    ap.add_argument("--disable_caching", action="store_true")
    args = ap.parse_args(argv)
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # launched by torchrun: one process per GPU, row-sharded indexes
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dataset = _load_from_disk(args.dataset)
    with open(args.config, "rt") as file:
        config = json.load(file)
    format_kwargs = config.pop("format", {})
    if format_kwargs:
        dataset.set_format(**format_kwargs)
    dataset_search(dataset, args.k, metric_save_path=Path(args.metrics) if args.metrics else None, **config)


if __name__ == "__main__":
    main()
