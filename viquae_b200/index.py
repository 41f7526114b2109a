"""`B200FlatIndex` — the engine's index object behind the HF-datasets index protocol.

Mirrors ``datasets.search.FaissIndex`` (datasets/search.py:233-414) for the one configuration the
reference uses on this path: ``faiss.IndexFlat(d, METRIC_INNER_PRODUCT)``, optionally preceded by an
L2 normalisation (``string_factory="L2norm,Flat"``, meerqat/ir/search.py:230-245).  It subclasses
``datasets.search.BaseIndex`` so ``dataset._indexes[name] = B200FlatIndex(...)`` makes
``Dataset.search_batch(name, queries, k)`` (datasets/search.py:713-733) dispatch to the GPU engine.

Host code only: every vector operation is a call into libviquae_b200.so (see ``_lib``); PyTorch
is used for device memory, streams and ``torch.distributed`` plumbing, never for arithmetic.
"""
from __future__ import annotations

import json
import os
from typing import NamedTuple, Optional

import numpy as np

from . import _lib

try:  # the reference depends on HF datasets (requirements.txt:1); keep working without it
    from datasets.search import BaseIndex, BatchedSearchResults, SearchResults
except Exception:  # pragma: no cover - exercised only where `datasets` is absent
    class SearchResults(NamedTuple):
        scores: list
        indices: list

    class BatchedSearchResults(NamedTuple):
        total_scores: list
        total_indices: list

    class BaseIndex:  # minimal stand-in with the same abstract surface
        def search(self, query, k: int = 10, **kwargs):
            raise NotImplementedError

        def search_batch(self, queries, k: int = 10, **kwargs):
            raise NotImplementedError

        def save(self, file):
            raise NotImplementedError

        @classmethod
        def load(cls, file):
            raise NotImplementedError


METRIC_INNER_PRODUCT = 0  # faiss.METRIC_INNER_PRODUCT; `"metric_type": 0` in the shipped configs
METRIC_L2 = 1
_DTYPES = {"f32": _lib.F32, "fp32": _lib.F32, "float32": _lib.F32, "bf16": _lib.BF16, "bfloat16": _lib.BF16,
           "f16": _lib.F16, "fp16": _lib.F16, "float16": _lib.F16, "f32x3": _lib.F32X3}
_WIDE = (_lib.F32, _lib.F32X3)  # storage types that hold full fp32 values
_VARIANTS = {None: _lib.VARIANT_AUTO, "auto": _lib.VARIANT_AUTO, "tc_cg1": _lib.VARIANT_TC_CG1,
             "tc_cg2": _lib.VARIANT_TC_CG2}


def parse_string_factory(string_factory: Optional[str], metric_type: Optional[int]) -> bool:
    """Returns do_L2norm for the factory strings the reference uses; raises on anything the engine
    does not serve.  ``None``/``"Flat"`` -> plain IP; ``"L2norm,Flat"`` -> cosine via normalisation
    (meerqat/ir/search.py:230-233).  ``string_factory=None`` with ``metric_type=None`` would build an
    **L2** index in datasets (datasets/search.py:283-286) — refused rather than silently different."""
    if metric_type not in (None, METRIC_INNER_PRODUCT):
        raise ValueError(f"B200 index kind supports inner product only (metric_type=0), got {metric_type}")
    if string_factory is None:
        if metric_type is None:
            raise ValueError("string_factory=None and metric_type=None selects an L2 index in datasets/FAISS; "
                             "the B200 index kind is inner-product only: pass metric_type=0")
        return False
    parts = [p.strip() for p in string_factory.split(",") if p.strip()]
    if parts == ["Flat"]:
        return False
    if parts == ["L2norm", "Flat"]:
        return True
    raise ValueError(f"B200 index kind supports string_factory in {{None, 'Flat', 'L2norm,Flat'}}, got {string_factory!r}")


def _torch():
    import torch
    return torch


def _tensor_dtype_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    if t.dtype == torch.float16:
        return _lib.F16
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


class B200FlatIndex(BaseIndex):
    """Exact inner-product top-k index living in the HBM of one B200.

    Parameters mirror ``FaissIndex(device, string_factory, metric_type)`` plus the storage ``dtype``
    ("bf16" default: tensor-core path; "f16"; "f32": exact fixed-order mode; "f32x3": fp32 values split into three
    bf16 terms and contracted on the tensor cores with fp32 accumulation — fp32-faithful scores, 12 bytes/element).
    """

    def __init__(self, d: Optional[int] = None, dtype: str = "bf16", device: int = 0, string_factory: Optional[str] = "Flat",
                 metric_type: Optional[int] = METRIC_INNER_PRODUCT, capacity: int = 0, variant: Optional[str] = None,
                 id_offset: int = 0):
        self.do_L2norm = parse_string_factory(string_factory, metric_type)
        self.string_factory = string_factory
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {dtype!r}")
        self.dtype = dtype
        self.device = int(device) if device is not None else 0
        if self.device < 0:
            raise ValueError("device must be a CUDA ordinal >= 0 (row-sharding over several GPUs: see viquae_b200.sharded)")
        self.variant = variant
        self.id_offset = int(id_offset)
        self._capacity = int(capacity)
        self._h = None
        self._ws = None  # torch uint8 workspace, grown on demand
        self.d = None
        if d is not None:
            self._create(int(d))

    # ------------------------------------------------------------------ lifecycle
    def _create(self, d: int):
        lib = _lib.load()
        if lib.vq_device_count() <= 0:
            raise RuntimeError("viquae_b200: no CUDA device visible; the engine has no CPU fallback")
        h = _lib.c_void_p()
        _lib.check(lib.vq_index_create(_lib.ctypes.byref(h), d, _DTYPES[self.dtype], self.device, self._capacity))
        self._h = h
        self.d = d
        if self.id_offset:
            _lib.check(lib.vq_index_set_id_offset(h, self.id_offset))
        if self.variant is not None:
            self.set_variant(self.variant)

    def close(self):
        if self._h is not None:
            _lib.load().vq_index_destroy(self._h)
            self._h = None
        self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # HF datasets deep-copies Dataset._indexes on every formatting/transform call; the index is a handle on
    # device memory, so copies share it (copying 18 GB of HBM per `with_format` is never what the caller wants)
    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self

    def __reduce__(self):
        raise TypeError("B200FlatIndex holds device memory and cannot be pickled: use save()/load()")

    def set_variant(self, variant: Optional[str]):
        if variant not in _VARIANTS:
            raise ValueError(f"variant must be one of {list(_VARIANTS)}")
        self.variant = variant
        if self._h is not None:
            _lib.check(_lib.load().vq_index_set_variant(self._h, _VARIANTS[variant]))

    def set_id_offset(self, off: int):
        self.id_offset = int(off)
        if self._h is not None:
            _lib.check(_lib.load().vq_index_set_id_offset(self._h, self.id_offset))

    @property
    def ntotal(self) -> int:
        return 0 if self._h is None else int(_lib.load().vq_index_ntotal(self._h))

    def __len__(self):
        return self.ntotal

    # ------------------------------------------------------------------ ingest
    def add(self, x, l2norm: Optional[bool] = None):
        """faiss ``index.add(x)``: x is a [n, d] numpy array (host) or torch tensor (host or cuda)."""
        torch = _torch()
        l2 = self.do_L2norm if l2norm is None else bool(l2norm)
        if isinstance(x, torch.Tensor):
            if x.dim() != 2:
                raise ValueError("Shape of vectors must be 2D")
            x = x.contiguous()
            if self._h is None:
                self._create(x.shape[1])
            if x.shape[1] != self.d:
                raise ValueError(f"dimension mismatch: index d={self.d}, vectors d={x.shape[1]}")
            if x.is_cuda:
                if x.device.index != self.device:
                    x = x.to(f"cuda:{self.device}")
                stream = torch.cuda.current_stream(self.device).cuda_stream
                _lib.check(_lib.load().vq_index_add(self._h, x.data_ptr(), x.shape[0], _tensor_dtype_code(x), int(l2), stream))
                torch.cuda.current_stream(self.device).synchronize()  # x may be freed by the caller
            else:
                _lib.check(_lib.load().vq_index_add(self._h, x.data_ptr(), x.shape[0], _tensor_dtype_code(x), int(l2), None))
            return
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError("Shape of vectors must be 2D")
        if self._h is None:
            self._create(x.shape[1])
        if x.shape[1] != self.d:
            raise ValueError(f"dimension mismatch: index d={self.d}, vectors d={x.shape[1]}")
        _lib.check(_lib.load().vq_index_add(self._h, x.ctypes.data, x.shape[0], _lib.F32, int(l2), None))

    def add_vectors(self, vectors, column: Optional[str] = None, batch_size: int = 65536, train_size=None,
                    faiss_verbose=None):
        """``FaissIndex.add_vectors`` (datasets/search.py:255-313): ``vectors`` is an array or a Dataset
        whose ``column`` holds the embeddings.  Rows go to the device in ``batch_size`` slices
        (the reference uses 1000-row numpy batches; the slice size only affects host staging)."""
        if column is not None and not isinstance(vectors, np.ndarray):
            if hasattr(vectors, "with_format"):  # HF Dataset: zero-copy Arrow chunks where possible
                from .ingest import arrow_column_blocks
                for block in arrow_column_blocks(vectors, column, max_rows=batch_size):
                    self.add(block)
                return
            n = len(vectors)
            for i in range(0, n, batch_size):
                batch = vectors[i:i + batch_size][column]
                self.add(np.asarray(batch, dtype=np.float32))
            return
        n = len(vectors)
        for i in range(0, n, batch_size):
            self.add(vectors[i:i + batch_size])

    # ------------------------------------------------------------------ search
    def _workspace(self, nbytes: int):
        torch = _torch()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def search_tensors(self, queries, k: int = 100, l2norm: Optional[bool] = None, out=None):
        """Device-resident search: ``queries`` a cuda tensor [nq, d] (f32/bf16/f16) -> (scores f32 [nq, k],
        ids int64 [nq, k]) cuda tensors (written into ``out=(scores, ids)`` when given).  Asynchronous on the
        current stream."""
        torch = _torch()
        if self._h is None:
            raise RuntimeError("empty index: add vectors first")
        if queries.dim() != 2:
            raise ValueError("Shape of query must be 2D")
        if queries.shape[1] != self.d:
            raise ValueError(f"dimension mismatch: index d={self.d}, queries d={queries.shape[1]}")
        if not queries.is_cuda or queries.device.index != self.device:
            raise ValueError(f"search_tensors needs queries on cuda:{self.device} (the index's device), got {queries.device}; "
                             "host arrays go through search_batch")
        q = queries.contiguous()
        nq = q.shape[0]
        dev = f"cuda:{self.device}"
        if out is not None:
            scores, ids = out
            assert scores.shape == (nq, k) and ids.shape == (nq, k) and scores.is_contiguous() and ids.is_contiguous()
        else:
            scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
            ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq == 0:
            return scores, ids
        lib = _lib.load()
        nbytes = lib.vq_index_workspace_bytes(self._h, nq, k)
        ws = self._workspace(nbytes)
        l2 = self.do_L2norm if l2norm is None else bool(l2norm)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.vq_index_search(self._h, q.data_ptr(), nq, _tensor_dtype_code(q), int(l2), k, scores.data_ptr(),
                                       ids.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return scores, ids

    def capture_search(self, nq: int, k: int = 100, query_dtype=None, l2norm: Optional[bool] = None) -> "CapturedSearch":
        """Records one ``search_tensors`` call of a fixed shape into a CUDA graph (every launch of the search is
        asynchronous on the caller's stream and touches only preallocated memory, so the whole multi-pass
        sequence replays as one submission).  For the latency regime (BASELINE config 5: batch 1-32 per call,
        interact/system.py's one-question loop).  The index must not grow while the graph is alive."""
        return CapturedSearch(self, nq, k, query_dtype, l2norm)

    def search_batch(self, queries, k: int = 10, **kwargs) -> BatchedSearchResults:
        """``FaissIndex.search_batch`` (datasets/search.py:369-385): numpy float32 [nq, d] in,
        (float32 [nq, k], int64 [nq, k]) numpy out through the host-buffer C-ABI call."""
        torch = _torch()
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            s, i = self.search_tensors(queries, k)
            return BatchedSearchResults(s, i)
        queries = np.asarray(queries, dtype=np.float32)
        if len(queries.shape) != 2:
            raise ValueError("Shape of query must be 2D")  # datasets/search.py:380-381
        if not queries.flags.c_contiguous:
            queries = np.asarray(queries, order="C")  # datasets/search.py:382-383
        if self._h is None:
            raise RuntimeError("empty index: add vectors first")
        if queries.shape[1] != self.d:
            raise ValueError(f"dimension mismatch: index d={self.d}, queries d={queries.shape[1]}")
        nq = queries.shape[0]
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        _lib.check(_lib.load().vq_index_search_host(self._h, queries.ctypes.data, nq, int(self.do_L2norm), k,
                                                    scores.ctypes.data, ids.ctypes.data))
        return BatchedSearchResults(scores, ids.astype(int))

    def search(self, query, k: int = 10, **kwargs) -> SearchResults:
        """``FaissIndex.search`` (datasets/search.py:349-367): one query vector."""
        query = np.asarray(query, dtype=np.float32)
        if len(query.shape) != 1 and (len(query.shape) != 2 or query.shape[0] != 1):
            raise ValueError("Shape of query is incorrect, it has to be either a 1D array or 2D (1, N)")
        res = self.search_batch(query.reshape(1, -1), k)
        return SearchResults(res.total_scores[0], res.total_indices[0])

    # ------------------------------------------------------------------ persistence
    def reconstruct_n(self, start: int = 0, n: Optional[int] = None, dtype: str = "f32") -> np.ndarray:
        n = self.ntotal - start if n is None else n
        code = _DTYPES[dtype]
        if code == _lib.F32X3:
            raise ValueError("reconstruct_n: 'f32x3' is a storage layout, ask for 'f32', 'bf16' or 'f16' rows")
        out = np.empty((n, self.d), dtype=np.float32 if code == _lib.F32 else np.uint16)
        if n:
            _lib.check(_lib.load().vq_index_copy_rows(self._h, start, n, out.ctypes.data, code, None))
        return out

    def save(self, file, storage_options=None):
        """Index persistence (``save_path`` in the reference configs, meerqat/ir/search.py:246-248):
        ``<file>`` = JSON header, ``<file>.npy`` = rows in the storage dtype (uint16 bit patterns for
        bf16/f16).  Rows are stored after ingest-time normalisation, so loading never re-normalises."""
        file = os.fspath(file)
        code = _DTYPES[self.dtype]
        rows = self.reconstruct_n(0, None, "f32" if code in _WIDE else self.dtype)
        np.save(file + ".npy", rows, allow_pickle=False)
        with open(file, "wt") as f:
            json.dump({"format": "viquae_b200.flat_ip.v1", "d": self.d, "ntotal": self.ntotal, "dtype": self.dtype,
                       "string_factory": self.string_factory, "id_offset": self.id_offset}, f)

    @classmethod
    def load(cls, file, device: int = 0, storage_options=None, variant: Optional[str] = None, dtype: str = "bf16") -> "B200FlatIndex":
        """Loads an index written by :meth:`save`, or a FAISS "Flat" / "L2norm,Flat" file written by the
        reference's ``save_path`` (``faiss.write_index``; recognised by its fourcc, stored as ``dtype``)."""
        file = os.fspath(file)
        with open(file, "rb") as f:
            magic = f.read(4)
        if magic in (b"IxFI", b"IxF2", b"IxFl", b"IxPT"):
            from .faiss_io import load_faiss_flat as _load
            return _load(file, dtype=dtype, device=device, variant=variant)
        with open(file, "rt") as f:
            hdr = json.load(f)
        if hdr.get("format") != "viquae_b200.flat_ip.v1":
            raise ValueError(f"{file}: not a viquae_b200 index file")
        rows = np.load(file + ".npy", allow_pickle=False)
        if rows.shape != (hdr["ntotal"], hdr["d"]):
            raise ValueError(f"{file}.npy: shape {rows.shape} does not match header ({hdr['ntotal']}, {hdr['d']})")
        idx = cls(d=hdr["d"], dtype=hdr["dtype"], device=device, string_factory=hdr["string_factory"],
                  capacity=hdr["ntotal"], variant=variant, id_offset=hdr.get("id_offset", 0))
        code = _DTYPES[hdr["dtype"]]
        if code == _lib.F32X3:
            code = _lib.F32  # saved as the exact fp32 values, split again on ingest
        if hdr["ntotal"]:
            rows = np.ascontiguousarray(rows)
            _lib.check(_lib.load().vq_index_add(idx._h, rows.ctypes.data, rows.shape[0], code, 0, None))
        return idx


class CapturedSearch:
    """A fixed-shape search held as a CUDA graph: write ``queries`` (or pass them to :meth:`__call__`), replay,
    read ``scores`` / ``ids``.  All three are static device tensors owned by this object."""

    def __init__(self, index: B200FlatIndex, nq: int, k: int, query_dtype=None, l2norm: Optional[bool] = None):
        torch = _torch()
        if index._h is None or index.ntotal == 0:
            raise RuntimeError("empty index: add vectors first")
        if nq < 1:
            raise ValueError("nq must be >= 1")
        self.index, self.nq, self.k = index, nq, k
        self.ntotal = index.ntotal
        dev = f"cuda:{index.device}"
        self.queries = torch.zeros((nq, index.d), dtype=query_dtype or torch.float32, device=dev)
        self.scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        self.ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        # own workspace: the index's shared scratch buffer may be reallocated by a later, larger search
        self._ws = torch.empty((_lib.load().vq_index_workspace_bytes(index._h, nq, k),), dtype=torch.uint8, device=dev)
        self._l2 = int(index.do_L2norm if l2norm is None else bool(l2norm))
        # capture on the INDEX's device, whatever the caller's current device is: torch.cuda.graph records the
        # current device's capture stream, and launches made on another device's stream would not be captured
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(index.device))
            with torch.cuda.stream(side):
                self._launch(side)  # warm-up outside capture: one-time function attributes, lazy module load
            torch.cuda.current_stream(index.device).wait_stream(side)
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch(torch.cuda.current_stream(index.device))

    def _launch(self, stream):
        q = self.queries
        _lib.check(_lib.load().vq_index_search(self.index._h, q.data_ptr(), self.nq, _tensor_dtype_code(q), self._l2, self.k,
                                               self.scores.data_ptr(), self.ids.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                               stream.cuda_stream))

    def replay(self):
        if self.index._h is None or self.index.ntotal != self.ntotal:
            raise RuntimeError("the index changed after capture_search(); capture again")
        self.graph.replay()
        return self.scores, self.ids

    def __call__(self, queries):
        self.queries.copy_(queries, non_blocking=True)
        return self.replay()


def load_faiss_flat(file, dtype: str = "bf16", device: int = 0, variant: Optional[str] = None) -> B200FlatIndex:
    """``faiss.read_index`` for "Flat" / "L2norm,Flat" files (see :mod:`viquae_b200.faiss_io`)."""
    from .faiss_io import load_faiss_flat as _load
    return _load(file, dtype=dtype, device=device, variant=variant)


def merge_topk(scores, ids, k_out: Optional[int] = None):
    """Cross-shard merge (K4): cuda tensors scores f32 / ids i64 [n_lists, nq, k_in] -> ([nq, k_out], [nq, k_out])."""
    torch = _torch()
    n_lists, nq, k_in = scores.shape
    k_out = k_in if k_out is None else k_out
    scores = scores.contiguous()
    ids = ids.contiguous()
    out_s = torch.empty((nq, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k_out), dtype=torch.int64, device=scores.device)
    if nq:
        with torch.cuda.device(scores.device):
            stream = torch.cuda.current_stream(scores.device).cuda_stream
            _lib.check(_lib.load().vq_merge_topk(scores.data_ptr(), ids.data_ptr(), n_lists, nq, k_in, k_out, out_s.data_ptr(),
                                                 out_i.data_ptr(), stream))
    return out_s, out_i


def packed_result_buffer(nq: int, k: int, device, n_lists: int = 1):
    """One byte buffer per list holding scores (f32 [nq, k]) then ids (i64 [nq, k]), so that the candidates of a
    rank travel in ONE all-gather.  Returns (buffer uint8 [n_lists, nbytes], scores views, ids views)."""
    torch = _torch()
    s_bytes = (nq * k * 4 + 7) // 8 * 8
    nbytes = s_bytes + nq * k * 8
    buf = torch.empty((n_lists, nbytes), dtype=torch.uint8, device=device)
    scores = [buf[c, :nq * k * 4].view(torch.float32).view(nq, k) for c in range(n_lists)]
    ids = [buf[c, s_bytes:].view(torch.int64).view(nq, k) for c in range(n_lists)]
    return buf, scores, ids


def merge_topk_packed(buf, nq: int, k_in: int, k_out: Optional[int] = None):
    """K4 over a gathered packed buffer [n_lists, nbytes] (see ``packed_result_buffer``)."""
    torch = _torch()
    n_lists, nbytes = buf.shape
    k_out = k_in if k_out is None else k_out
    s_bytes = (nq * k_in * 4 + 7) // 8 * 8
    out_s = torch.empty((nq, k_out), dtype=torch.float32, device=buf.device)
    out_i = torch.empty((nq, k_out), dtype=torch.int64, device=buf.device)
    if nq:
        with torch.cuda.device(buf.device):
            stream = torch.cuda.current_stream(buf.device).cuda_stream
            _lib.check(_lib.load().vq_merge_topk_strided(buf.data_ptr(), buf.data_ptr() + s_bytes, nbytes // 4, nbytes // 8, n_lists, nq, k_in,
                                                         k_out, out_s.data_ptr(), out_i.data_ptr(), stream))
    return out_s, out_i
