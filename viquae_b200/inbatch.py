"""Rows next to the search path: where the embeddings come from and go to.

* :func:`inbatch_step` — the evaluation step of the reference's bi-encoder trainee
  (meerqat/train/trainee.py:436-458 ``BiEncoder.step`` + ``eval_step``, meerqat/train/metrics.py:22-37
  ``batch_retrieval``): all-gather of question / context embeddings and labels across ranks with the label shift,
  ``similarities = Q @ C.T``, ``log_softmax``, NLL loss, MRR and hits@1 of the labelled context — the similarity,
  log-sum-exp, loss term and rank come from ONE fused kernel (``vq_inbatch_logprobs``).  Forward only: training
  (autograd) stays with the reference's PyTorch code.
* :func:`mine_hard_negatives` — EXPERIMENTS.rst:78-103: search the training questions (64k-query batches run in the
  tensor-core-bound regime, BASELINE config 4), then split every question's retrieved ids into relevant ones and
  ``search_irrelevant_indices``.  The embeddings stay on the device from the encoder to the search (the reference goes
  GPU -> ``.cpu().numpy()`` -> Arrow -> list -> ``np.array``, meerqat/ir/embedding.py:223-246).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import _lib

IGNORE_INDEX = -100  # torch.nn.NLLLoss default, the reference's loss_fct.ignore_index


def _code(t):
    import torch
    return {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}[t.dtype]


def gather_global_batch(question_embeddings, context_embeddings, labels, ignore_index: int = IGNORE_INDEX, group=None):
    """trainee.py:438-452: all-gather the three tensors over the ranks and shift every rank's labels by the number of
    contexts of the ranks before it (labels index the CONCATENATED contexts).  Without a process group: unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return question_embeddings, context_embeddings, labels
    G, r = dist.get_world_size(group), dist.get_rank(group)
    n_ctx = context_embeddings.shape[0]
    shifted = torch.where(labels != ignore_index, labels + r * n_ctx, labels)
    out = []
    for t in (question_embeddings.contiguous(), context_embeddings.contiguous(), shifted.contiguous()):
        buf = torch.empty((G * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)
        out.append(buf)
    return tuple(out)


def inbatch_step(question_embeddings, context_embeddings, labels, ignore_index: int = IGNORE_INDEX, return_log_probs: bool = True,
                 group=None):
    """CUDA tensors: questions [N, d], contexts [N*M, d] (f32 / bf16 / f16), labels int64 [N] -> dict with ``loss``
    (NLLLoss 'mean' over the non-ignored rows), ``log_probs`` [N, N*M] float32 (optional), ``global_labels``, ``ranks``
    (1-based rank of the labelled context, 0 = ignored) and ``metrics`` = what ``batch_retrieval`` returns
    (``MRR@N*M`` and ``hits@1`` as SUMS over the batch, ``ignored_predictions``, ``batch_size``)."""
    import torch

    q, c, y = gather_global_batch(question_embeddings, context_embeddings, labels, ignore_index, group)
    if not q.is_cuda:
        raise ValueError("inbatch_step works on CUDA tensors (the engine has no CPU path)")
    q, c, y = q.contiguous(), c.to(q.dtype).contiguous(), y.to(torch.int64).contiguous()
    n_q, d = q.shape
    n_c = c.shape[0]
    dev = q.device
    log_probs = torch.empty((n_q, n_c), dtype=torch.float32, device=dev) if return_log_probs else None
    nll = torch.empty((n_q,), dtype=torch.float32, device=dev)
    rank = torch.empty((n_q,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().vq_inbatch_logprobs(q.data_ptr(), c.data_ptr(), _code(q), n_q, n_c, d, y.data_ptr(), ignore_index,
                                                   log_probs.data_ptr() if log_probs is not None else None, nll.data_ptr(), rank.data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream))
    ranks = rank.cpu().numpy()
    nll_h = nll.cpu().numpy()
    kept = ranks > 0
    n_kept = int(kept.sum())
    loss = float(nll_h[kept].astype(np.float64).sum() / n_kept) if n_kept else float("nan")
    metrics = {"MRR@N*M": float((1.0 / ranks[kept]).sum()) if n_kept else 0.0, "hits@1": int((ranks[kept] == 1).sum()),
               "ignored_predictions": int(n_q - n_kept), "batch_size": int(n_q)}
    return {"loss": loss, "log_probs": log_probs, "global_labels": y, "ranks": rank, "metrics": metrics}


def mine_hard_negatives(index, query_embeddings, relevant: Sequence[Sequence[int]], k: int = 100, fallback: Optional[Sequence[Sequence[int]]] = None,
                        batch_size: int = 65536):
    """EXPERIMENTS.rst:78-103 for dense retrieval: ``query_embeddings`` (a CUDA tensor straight from the encoder, [n, d])
    are searched in ``batch_size`` chunks against ``index`` (a ``B200FlatIndex`` or a sharded one); every retrieved id is
    marked relevant iff it is in ``relevant[i]`` (the ids ``find_relevant`` accepted for question i).  Returns two lists
    of lists like the snippet's ``search_provenance_indices`` (relevant retrieved ids, best first; ``fallback[i]`` when
    none was retrieved) and ``search_irrelevant_indices`` (the hard negatives, best first)."""
    import torch

    n = query_embeddings.shape[0]
    if len(relevant) != n:
        raise ValueError("one list of relevant ids per query")
    dev = query_embeddings.device
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(set(r)) for r in relevant], out=offsets[1:])
    flat = np.fromiter((i for r in relevant for i in sorted(set(int(x) for x in r))), dtype=np.int64, count=int(offsets[-1]))
    d_off = torch.from_numpy(offsets).to(dev)
    d_rel = torch.from_numpy(flat if len(flat) else np.zeros(1, np.int64)).to(dev)
    lib = _lib.load()
    provenance: List[List[int]] = []
    irrelevant: List[List[int]] = []
    for a in range(0, n, batch_size):
        b = min(n, a + batch_size)
        _, ids = index.search_tensors(query_embeddings[a:b].contiguous(), k)
        ids = ids.contiguous()
        mask = torch.empty((b - a, k), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.vq_mark_relevant(ids.data_ptr(), b - a, k, d_off[a:].data_ptr(), d_rel.data_ptr(), mask.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))
        ids_h, mask_h = ids.cpu().numpy(), mask.cpu().numpy().astype(bool)
        for j in range(b - a):
            valid = ids_h[j] >= 0
            rel_j = ids_h[j][valid & mask_h[j]].tolist()
            if not rel_j and fallback is not None:
                rel_j = list(fallback[a + j])
            provenance.append(rel_j)
            irrelevant.append(ids_h[j][valid & ~mask_h[j]].tolist())
    return provenance, irrelevant
