"""Ingest helpers: the embedding column of an HF Dataset -> float32 row blocks for
``B200FlatIndex.add`` without going through Python lists.

The reference builds its FAISS index by iterating the dataset in 1000-row numpy batches
(datasets/search.py:309-313: ``vectors[i:i+batch_size][column]``), i.e. Arrow ``list<float>`` ->
Python objects -> numpy for every batch.  The embeddings are written as fixed-length float lists
(ir/embedding.py:238), so the Arrow values buffer of each chunk already IS the row-major matrix:
``arrow_column_blocks`` exposes it zero-copy and only falls back to the formatted read when a chunk
has ragged rows, nulls or a non-float32 value type.
"""
from __future__ import annotations

from typing import Iterator

import numpy as np


def _chunk_to_matrix(chunk):
    """One pyarrow (Large)ListArray / FixedSizeListArray chunk -> float32 [n, d] view, or None if not uniform."""
    import pyarrow as pa

    if chunk.null_count:
        return None
    t = chunk.type
    if pa.types.is_fixed_size_list(t):
        d = t.list_size
        values = chunk.flatten()  # respects the chunk's offset/length
    elif pa.types.is_list(t) or pa.types.is_large_list(t):
        offsets = chunk.offsets.to_numpy()
        n = len(chunk)
        if n == 0:
            return np.zeros((0, 0), np.float32)
        d = int(offsets[1] - offsets[0])
        if d <= 0 or not np.array_equal(offsets, offsets[0] + d * np.arange(n + 1, dtype=offsets.dtype)):
            return None  # ragged
        values = chunk.flatten()
    else:
        return None
    if values.null_count or not pa.types.is_floating(values.type):
        return None
    flat = values.to_numpy(zero_copy_only=False)
    if flat.dtype != np.float32:
        flat = flat.astype(np.float32)
    return flat.reshape(len(chunk), d)


def arrow_column_blocks(dataset, column: str, max_rows: int = 262144, start: int = 0, stop: int = None) -> Iterator[np.ndarray]:
    """Yields float32 [n, d] blocks covering rows ``[start, stop)`` of ``dataset[column]`` in order (the whole column
    by default; a row shard of it for a sharded index).  Zero-copy from the Arrow chunks where the column is a uniform
    float list; otherwise falls back to ``dataset.with_format('numpy')`` slices."""
    n = len(dataset)
    stop = n if stop is None else min(int(stop), n)
    start = max(0, int(start))
    table = getattr(dataset, "data", None)
    indices = getattr(dataset, "_indices", None)
    if table is not None and indices is None and hasattr(table, "column"):
        col = table.column(column)
        ok = True
        blocks = []
        row0 = 0
        for chunk in col.chunks:
            a, b = max(start, row0), min(stop, row0 + len(chunk))
            if b > a:
                m = _chunk_to_matrix(chunk.slice(a - row0, b - a))
                if m is None:
                    ok = False
                    break
                blocks.append(m)
            row0 += len(chunk)
        if ok:
            for m in blocks:
                for i in range(0, len(m), max_rows):
                    yield m[i:i + max_rows]
            return
    ds = dataset.with_format("numpy", columns=[column]) if hasattr(dataset, "with_format") else dataset
    for i in range(start, stop, max_rows):
        yield np.asarray(ds[i:min(stop, i + max_rows)][column], dtype=np.float32)
