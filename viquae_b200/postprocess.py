"""Device-side run post-processing (K6): the per-(query, hit) loop of ``Searcher.__call__``
(meerqat/ir/search.py:413-440) as one kernel launch per index and batch.

``CSRMapping`` holds the article -> passages mapping (the reference's ``index_mapping`` JSON dict,
meerqat/data/loading.py:214-232) as CSR arrays on the device; ``expand_hits`` applies it to the
``(scores, ids)`` a search returned and yields the run of every query in dict insertion order.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import _lib


class CSRMapping:
    def __init__(self, index_mapping: Dict[int, List[int]], n_rows: Optional[int] = None, device: int = 0):
        import torch

        n = (max(index_mapping) + 1 if index_mapping else 0) if n_rows is None else int(n_rows)
        deg = np.zeros(n, dtype=np.int64)
        for row, targets in index_mapping.items():
            if 0 <= row < n:
                deg[row] = len(targets)
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(deg, out=offsets[1:])
        flat = np.empty(int(offsets[-1]), dtype=np.int64)
        for row, targets in index_mapping.items():
            if 0 <= row < n:
                flat[offsets[row]:offsets[row + 1]] = targets
        self.n_rows = n
        self.max_degree = int(deg.max()) if n else 0
        dev = torch.device(f"cuda:{device}")
        self.offsets = torch.from_numpy(offsets).to(dev)
        self.targets = torch.from_numpy(flat).to(dev)


def expand_hits(scores, ids, mapping: Optional[CSRMapping], many2one: Optional[str], k: int):
    """scores f32 / ids i64 cuda tensors [nq, k_in] -> (run_scores f64 [nq, cap], run_ids i64 [nq, cap], counts i32 [nq])
    on the same device, documents in dict insertion order (see csrc/vq_post.cu)."""
    import torch

    if many2one not in (None, "max"):
        raise ValueError(f"Invalid value for many2one: '{many2one}'. Choose from {{None, 'max'}}")
    scores = scores.contiguous()
    ids = ids.contiguous()
    nq, k_in = scores.shape
    cap = k + (max(mapping.max_degree, 1) if mapping is not None else 1) - 1
    cap = max(cap, 1)
    dev = scores.device
    out_s = torch.empty((nq, cap), dtype=torch.float64, device=dev)
    out_i = torch.empty((nq, cap), dtype=torch.int64, device=dev)
    out_c = torch.zeros((nq,), dtype=torch.int32, device=dev)
    flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    if nq:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().vq_expand_hits(
                scores.data_ptr(), ids.data_ptr(), nq, k_in,
                mapping.offsets.data_ptr() if mapping is not None else None,
                mapping.targets.data_ptr() if mapping is not None else None,
                1 if many2one == "max" else 0, k, cap, out_s.data_ptr(), out_i.data_ptr(), out_c.data_ptr(), flag.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream))
        if int(flag.item()):
            raise RuntimeError("vq_expand_hits: output capacity exceeded (mapping fan-out larger than declared)")
    return out_s, out_i, out_c
