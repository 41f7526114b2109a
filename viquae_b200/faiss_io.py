"""Interchange with FAISS index files for the one index family the reference writes on this path.

The reference persists its dense indexes with ``Dataset.save_faiss_index`` -> ``faiss.write_index``
(meerqat/ir/search.py:246-248, datasets/search.py:387-397) and the only factories it uses are ``"Flat"`` and
``"L2norm,Flat"`` with ``metric_type=0`` (experiments/**/config.json), i.e. on disk

* ``IndexFlat``           — fourcc ``IxFI`` (inner product) / ``IxF2`` (L2) / ``IxFl`` (other metric)
* ``IndexPreTransform``   — fourcc ``IxPT``, a chain holding one ``NormalizationTransform`` (``VNrm``), then ``IndexFlat``

FAISS is a third-party dependency that is absent from /root/reference and from this image (requirements.txt:14
pins ``faiss-gpu>=1.7.1``).  The byte layout below restates the published serialiser of faiss 1.7.x
(``faiss/impl/index_write.cpp``: ``write_index_header``, ``WRITEXBVECTOR``, ``write_VectorTransform``); all
fields little-endian, no padding:

    index header : d int32 | ntotal int64 | dummy int64 (1<<20) | dummy int64 | is_trained uint8 | metric_type int32
                   [| metric_arg float32   only when metric_type > 1]
    IndexFlat    : fourcc | header | n_words uint64 (= ntotal*d, vector length in 4-byte units) | float32[ntotal*d]
    VNrm         : fourcc | norm float32 | d_in int32 | d_out int32 | is_trained uint8
    IxPT         : fourcc | header | n_transforms int32 | transforms... | sub-index

PARITY UNPINNED: there is no FAISS here to produce or read a real file, so the tests only pin this module
against hand-assembled byte strings of the layout above and against its own round trip.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import BinaryIO, Optional

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
_DUMMY = 1 << 20


@dataclass
class FlatFile:
    """What an ``IndexFlat`` (optionally behind ``L2norm``) file holds."""
    rows: np.ndarray          # [ntotal, d] float32, exactly the stored values (already normalised when l2norm is set)
    metric_type: int          # faiss MetricType; 0 = inner product
    l2norm: bool              # an L2 NormalizationTransform precedes the flat index ("L2norm,Flat")

    @property
    def string_factory(self) -> str:
        return "L2norm,Flat" if self.l2norm else "Flat"


def _read(f: BinaryIO, fmt: str):
    size = struct.calcsize(fmt)
    buf = f.read(size)
    if len(buf) != size:
        raise ValueError("truncated FAISS index file")
    return struct.unpack(fmt, buf)


def _read_header(f: BinaryIO):
    d, ntotal, _, _, is_trained, metric_type = _read(f, "<iqqqBi")
    if metric_type > 1:
        _read(f, "<f")
    if d <= 0 or ntotal < 0:
        raise ValueError(f"corrupt FAISS index header (d={d}, ntotal={ntotal})")
    return d, ntotal, bool(is_trained), metric_type


def _write_header(f: BinaryIO, d: int, ntotal: int, metric_type: int):
    f.write(struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1, metric_type))
    if metric_type > 1:
        f.write(struct.pack("<f", 0.0))


def _read_flat(f: BinaryIO, fourcc: bytes) -> tuple:
    d, ntotal, _, metric_type = _read_header(f)
    expected = {b"IxFI": METRIC_INNER_PRODUCT, b"IxF2": METRIC_L2}.get(fourcc)
    if expected is not None and metric_type != expected:
        raise ValueError(f"FAISS file: fourcc {fourcc!r} disagrees with metric_type {metric_type}")
    (n_words,) = _read(f, "<Q")
    if n_words != ntotal * d:
        raise ValueError(f"FAISS IndexFlat: vector of {n_words} floats for ntotal={ntotal}, d={d}")
    rows = np.fromfile(f, dtype="<f4", count=n_words) if hasattr(f, "fileno") else np.frombuffer(f.read(4 * n_words), dtype="<f4")
    if rows.size != n_words:
        raise ValueError("truncated FAISS index file")
    return rows.reshape(ntotal, d).astype(np.float32, copy=False), metric_type


def read_flat_index(file) -> FlatFile:
    """Parse a file written by ``faiss.write_index`` for ``"Flat"`` / ``"L2norm,Flat"``.

    Anything else (IVF, PQ, HNSW, other transforms) raises ``NotImplementedError``: those are approximate
    indexes the reference never builds on this path (SURVEY §8 out of scope)."""
    with open(os.fspath(file), "rb") as f:
        fourcc = f.read(4)
        if fourcc in (b"IxFI", b"IxF2", b"IxFl"):
            rows, metric = _read_flat(f, fourcc)
            return FlatFile(rows, metric, False)
        if fourcc == b"IxPT":
            d, ntotal, _, metric = _read_header(f)
            (nt,) = _read(f, "<i")
            l2norm = False
            for _ in range(nt):
                vt = f.read(4)
                if vt != b"VNrm":
                    raise NotImplementedError(f"FAISS vector transform {vt!r} is not used by the reference's dense indexes")
                norm, d_in, d_out, _ = _read(f, "<fiiB")
                if norm != 2.0 or d_in != d or d_out != d:
                    raise NotImplementedError(f"NormalizationTransform(norm={norm}, {d_in}->{d_out}) is not L2norm over d={d}")
                l2norm = True
            sub = f.read(4)
            if sub not in (b"IxFI", b"IxF2", b"IxFl"):
                raise NotImplementedError(f"FAISS sub-index {sub!r} is not IndexFlat")
            rows, sub_metric = _read_flat(f, sub)
            if rows.shape != (ntotal, d) or sub_metric != metric:
                raise ValueError("FAISS IndexPreTransform header disagrees with its flat sub-index")
            return FlatFile(rows, metric, l2norm)
        raise NotImplementedError(f"FAISS index type {fourcc!r} is not IndexFlat / IndexPreTransform(L2norm, IndexFlat)")


def write_flat_index(file, rows: np.ndarray, metric_type: int = METRIC_INNER_PRODUCT, l2norm: bool = False) -> None:
    """Write ``rows`` ([ntotal, d] float32, already normalised when ``l2norm``) as FAISS would for
    ``index_factory(d, "Flat" | "L2norm,Flat", metric_type)`` after ``add``."""
    rows = np.ascontiguousarray(rows, dtype="<f4")
    if rows.ndim != 2:
        raise ValueError("rows must be [ntotal, d]")
    ntotal, d = rows.shape
    flat = {METRIC_INNER_PRODUCT: b"IxFI", METRIC_L2: b"IxF2"}.get(metric_type, b"IxFl")
    with open(os.fspath(file), "wb") as f:
        if l2norm:
            f.write(b"IxPT")
            _write_header(f, d, ntotal, metric_type)
            f.write(struct.pack("<i", 1))
            f.write(b"VNrm" + struct.pack("<fiiB", 2.0, d, d, 1))
        f.write(flat)
        _write_header(f, d, ntotal, metric_type)
        f.write(struct.pack("<Q", ntotal * d))
        rows.tofile(f)


def load_faiss_flat(file, dtype: str = "bf16", device: int = 0, variant: Optional[str] = None):
    """Build a :class:`viquae_b200.index.B200FlatIndex` from a FAISS flat inner-product file.  Stored rows are
    ingested as they are: ``IndexPreTransform.add`` runs the chain before the flat ``add``, so the rows of an
    ``"L2norm,Flat"`` file are already unit length and only the queries still need normalising."""
    from .index import B200FlatIndex
    ff = read_flat_index(file)
    if ff.metric_type != METRIC_INNER_PRODUCT:
        raise ValueError(f"{file}: metric_type {ff.metric_type} — this engine is inner-product only")
    idx = B200FlatIndex(d=ff.rows.shape[1], dtype=dtype, device=device, string_factory=ff.string_factory,
                        capacity=ff.rows.shape[0], variant=variant)
    if ff.rows.shape[0]:
        idx.add(ff.rows, l2norm=False)
    return idx


def save_faiss_flat(index, file) -> None:
    """Write a :class:`B200FlatIndex` as a FAISS file (fp32 upcast of the stored rows)."""
    rows = index.reconstruct_n(0, None, "f32")
    write_flat_index(file, rows, METRIC_INNER_PRODUCT, l2norm=bool(index.do_L2norm))
