"""Row-sharded multi-GPU index: one process per GPU, corpus rows split contiguously over ranks.

The reference reaches several GPUs only through FAISS *replicas* (``device=-1`` ->
``index_cpu_to_all_gpus``, datasets/search.py:329-340): every GPU holds the whole corpus.  Here
rank r owns rows ``[r*N/G, (r+1)*N/G)``, queries are replicated, every rank searches
its shard, the per-rank ``[nq, k]`` candidates are exchanged with ONE all-gather (NCCL over
NVLink / NVSwitch; gloo in the CPU tests) and each rank merges ``[G, nq, k]`` with the K4 kernel:
top-k of a union == top-k of the per-shard top-k's.

`local_search` / `merge` are injectable so that the collective plumbing can be exercised on CPU with
the ``gloo`` backend (tests/test_dist_gloo.py injects the oracle there; the product defaults are the
CUDA engine and nothing else).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def resolve_devices(device, world_size: int, rank: int, n_visible: int):
    """What ``device`` of an index config means for this process (``add_faiss_index(device=...)``,
    datasets/search.py:316-347: ``None`` CPU, int >= 0 that GPU, negative int ALL GPUs, list those GPUs).

    Returns ``(local_device, sharded)``.  The reference replicates the index over the GPUs from ONE process
    (``index_cpu_to_all_gpus``); this engine row-shards it over one process per GPU (``torchrun``), so a multi-GPU
    ``device`` needs an initialised process group whose size is the number of GPUs asked for:
      * ``None`` / int >= 0 -> that one GPU (``None`` = 0, or this rank's own GPU under ``torchrun``), not sharded;
      * ``-1`` (any negative) -> every rank's own GPU ``rank % n_visible``, sharded over the whole group;
      * list ``[g0, g1, ...]`` -> rank r owns GPU ``g_r``; the list length must equal the group size."""
    if device is None:
        return (rank % max(1, n_visible) if world_size > 1 else 0), False
    if isinstance(device, (list, tuple)):
        devs = [int(g) for g in device]
        if not devs or any(g < 0 for g in devs) or len(set(devs)) != len(devs):
            raise ValueError(f"device list must hold distinct GPU ordinals >= 0, got {device!r}")
        if len(devs) == 1:
            return devs[0], False
        if world_size != len(devs):
            raise RuntimeError(f"device={devs} asks for a corpus row-sharded over {len(devs)} GPUs: launch one process per GPU "
                               f"(python -m torch.distributed.run --nproc-per-node {len(devs)} ...); this process group has {world_size} rank(s)")
        return devs[rank], True
    device = int(device)
    if device >= 0:
        return device, False
    if world_size == 1:
        if n_visible > 1:
            raise RuntimeError(f"device={device} asks for all {n_visible} GPUs: launch one process per GPU "
                               f"(python -m torch.distributed.run --nproc-per-node {n_visible} ...) so that the corpus can be row-sharded")
        return 0, False
    return rank % max(1, n_visible), True


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: sizes differ by at most one row, earlier ranks take the extra."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class PeerExchange:
    """One NVLink peer-memory exchange window per rank (``vq_xchg_*``, csrc/vq_xchg.cu): the IPC handles travel once
    through ``torch.distributed.all_gather_object``; after that ``publish`` + ``merge_topk`` / ``gather`` are plain
    kernel launches on the caller's stream (no NCCL call, no host synchronisation, CUDA-graph capturable)."""

    def __init__(self, device: int, rank: int, world_size: int, group=None, slot_bytes: int = 16 << 20):
        import ctypes

        import torch.distributed as dist

        from . import _lib
        self._lib = _lib
        lib = _lib.load()
        self.device, self.rank, self.world_size = device, rank, world_size
        h = _lib.c_void_p()
        _lib.check(lib.vq_xchg_create(ctypes.byref(h), device, rank, world_size, slot_bytes))
        self._h = h
        self.slot_bytes = int(lib.vq_xchg_slot_bytes(h))
        buf = ctypes.create_string_buffer(64)
        _lib.check(lib.vq_xchg_export(h, buf))
        handles = [None] * world_size
        dist.all_gather_object(handles, bytes(buf.raw), group=group)
        for r, raw in enumerate(handles):
            if r != rank:
                _lib.check(lib.vq_xchg_attach(h, r, ctypes.create_string_buffer(raw, 64)))
        dist.barrier(group=group)  # nobody publishes before every window is mapped everywhere

    def publish(self, packed, stream):
        self._lib.check(self._lib.load().vq_xchg_publish(self._h, packed.data_ptr(), packed.numel() * packed.element_size(), stream))

    def merge_topk(self, nq, k_in, k_out, out_s, out_i, stream):
        self._lib.check(self._lib.load().vq_xchg_merge_topk(self._h, nq, k_in, k_out, out_s.data_ptr(), out_i.data_ptr(), stream))

    def gather(self, nbytes, dst, stream):
        self._lib.check(self._lib.load().vq_xchg_gather(self._h, nbytes, dst.data_ptr(), stream))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.load().vq_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedFlatIndex:
    """The local shard (a ``B200FlatIndex`` with ``id_offset`` = first global row) + the exchange step."""

    def __init__(self, local_index=None, *, rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None, transport: str = "peer"):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.local_index = local_index
        self._packed_ok = local_search is None and merge is None and local_index is not None
        if local_search is None:
            if local_index is None:
                raise ValueError("need a local index")
            local_search = local_index.search_tensors
        if merge is None:
            from .index import merge_topk
            merge = merge_topk
        self._local_search = local_search
        self._merge = merge
        # NVLink peer-memory exchange (results) and a second window for query slices; built lazily at the first search
        # that can use them, "nccl" transport on request or when the windows cannot be mapped (not one node / no IPC)
        import os
        self.transport = os.environ.get("VIQUAE_B200_TRANSPORT", transport)  # "peer" (default) | "nccl"
        self._xchg = None
        self._xchg_q = None
        self._xchg_failed = False

    PEER_SLOT_BYTES = 16 << 20  # per-rank window slot: 8,191 queries x top-100 x 12 B fits

    def _exchange(self, queries_window: bool = False):
        """The peer-memory exchange of this index (created collectively at first use), or None for the NCCL transport."""
        if self.transport != "peer" or self._xchg_failed or self.local_index is None:
            return None
        if self._xchg is None:
            try:
                dev = self.local_index.device
                self._xchg = PeerExchange(dev, self.rank, self.world_size, self.group, self.PEER_SLOT_BYTES)
                self._xchg_q = PeerExchange(dev, self.rank, self.world_size, self.group, self.PEER_SLOT_BYTES)
            except Exception as e:  # e.g. ranks on different nodes: every rank fails the same way and uses NCCL
                import warnings
                warnings.warn(f"viquae_b200: NVLink peer exchange unavailable ({e}); using the NCCL all-gather transport")
                self._xchg_failed = True
                return None
        return self._xchg_q if queries_window else self._xchg

    # above this many queries the candidates are exchanged by query slice (all-to-all), each rank merges nq/G
    # queries and the merged slices are all-gathered: G x less merge work and G x less traffic than gathering
    # every rank's [nq, k] on every rank (64k-query batches)
    SLICE_EXCHANGE_MIN_QUERIES = 8192

    def _search_by_query_slice(self, s_loc, i_loc, k):
        import torch

        G, nq = self.world_size, s_loc.shape[0]
        per = (nq + G - 1) // G
        pad = per * G - nq
        if pad:
            s_loc = torch.cat([s_loc, torch.full((pad, k), float("-inf"), dtype=s_loc.dtype, device=s_loc.device)])
            i_loc = torch.cat([i_loc, torch.full((pad, k), -1, dtype=i_loc.dtype, device=i_loc.device)])
        s_recv = torch.empty_like(s_loc)   # [G, per, k]: slice `rank` of every rank's candidates
        i_recv = torch.empty_like(i_loc)
        self.dist.all_to_all_single(s_recv, s_loc.contiguous(), group=self.group)
        self.dist.all_to_all_single(i_recv, i_loc.contiguous(), group=self.group)
        s_m, i_m = self._merge(s_recv.view(G, per, k), i_recv.view(G, per, k), k)
        s_all = torch.empty((G * per, k), dtype=s_m.dtype, device=s_m.device)
        i_all = torch.empty((G * per, k), dtype=i_m.dtype, device=i_m.device)
        self.dist.all_gather_into_tensor(s_all, s_m.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(i_all, i_m.contiguous(), group=self.group)
        return s_all[:nq], i_all[:nq]

    def search_tensors(self, queries, k: int = 100, l2norm: Optional[bool] = None):
        """queries: the SAME [nq, d] tensor on every rank (device of the local shard).
        Returns the global (scores [nq, k], ids [nq, k]) on every rank.  ``l2norm`` overrides the local index's
        query normalisation (``search_host`` normalises before the queries travel)."""
        import torch

        if l2norm is not None:
            local_search = lambda q, kk, out=None: self.local_index.search_tensors(q, kk, l2norm=l2norm, out=out)  # noqa: E731
        else:
            local_search = None

        if (self.world_size > 1 and self._packed_ok and queries.shape[0] < self.SLICE_EXCHANGE_MIN_QUERIES
                and self.dist.get_backend(self.group) == "nccl"):
            # scores + ids of this rank in one packed buffer
            from .index import merge_topk_packed, packed_result_buffer
            nq = queries.shape[0]
            mine, s_v, i_v = packed_result_buffer(nq, k, queries.device)
            self.local_index.search_tensors(queries, k, l2norm=l2norm, out=(s_v[0], i_v[0]))
            xchg = self._exchange() if (k <= 128 and mine.shape[1] <= self.PEER_SLOT_BYTES) else None
            if xchg is not None:
                # publish into this rank's NVLink window; the merge kernel waits for every rank's flag and pulls the
                # candidates straight from the peers' windows (exchange + merge = one kernel)
                out_s = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
                out_i = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
                with torch.cuda.device(queries.device):
                    stream = torch.cuda.current_stream(queries.device).cuda_stream
                    xchg.publish(mine, stream)
                    xchg.merge_topk(nq, k, k, out_s, out_i, stream)
                return out_s, out_i
            # NCCL transport: ONE all-gather of the packed buffers -> strided K4 merge
            gathered = torch.empty((self.world_size, mine.shape[1]), dtype=torch.uint8, device=queries.device)
            self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
            return merge_topk_packed(gathered, nq, k)
        s_loc, i_loc = (local_search or self._local_search)(queries, k)
        if self.world_size == 1:
            return s_loc, i_loc
        nq = s_loc.shape[0]
        if nq >= self.SLICE_EXCHANGE_MIN_QUERIES and self.dist.get_backend(self.group) == "nccl":
            return self._search_by_query_slice(s_loc, i_loc, k)
        s_all = torch.empty((self.world_size, nq, k), dtype=s_loc.dtype, device=s_loc.device)
        i_all = torch.empty((self.world_size, nq, k), dtype=i_loc.dtype, device=i_loc.device)
        # two small all-gathers (scores f32, ids i64): nq*k*12 B per rank, latency-bound on NVSwitch
        if self.dist.get_backend(self.group) == "nccl":
            self.dist.all_gather_into_tensor(s_all, s_loc.contiguous(), group=self.group)
            self.dist.all_gather_into_tensor(i_all, i_loc.contiguous(), group=self.group)
        else:  # gloo (CPU tests): same exchange through the list form
            self.dist.all_gather(list(s_all.unbind(0)), s_loc.contiguous(), group=self.group)
            self.dist.all_gather(list(i_all.unbind(0)), i_loc.contiguous(), group=self.group)
        return self._merge(s_all, i_all, k)

    def capture_search(self, nq: int, k: int = 100, query_dtype=None):
        """A fixed-shape sharded search (local passes + publish + pull-merge) recorded into ONE CUDA graph per rank: the
        latency regime at several GPUs (BASELINE config 5).  Needs the peer-memory transport (NCCL calls are not captured)."""
        return CapturedShardedSearch(self, nq, k, query_dtype)

    def _transport_dtype(self):
        """dtype the queries travel in between ranks: the index's own 16-bit type (rounding is what the local search
        would do anyway, and it is deterministic), full fp32 for the fp32 storage modes."""
        import torch

        name = getattr(self.local_index, "dtype", "f32")
        return {"bf16": torch.bfloat16, "f16": torch.float16, "fp16": torch.float16, "bfloat16": torch.bfloat16,
                "float16": torch.float16}.get(name, torch.float32)

    def search_host(self, queries, k: int = 100, out=None, result_ranks="all"):
        """Host entry point of the sharded engine: the SAME host float32 queries [nq, d] on every rank.

        Rank r copies only its slice ``[r*nq/G, (r+1)*nq/G)`` to the device, casts it to the transport dtype
        (``vq_prep_rows``) and one all-gather over NVLink rebuilds the batch on every GPU — instead of G full
        H2D copies through the host's PCIe root.  The merged result is copied back on ``result_ranks`` ("all" or a
        rank number; other ranks return ``None``) into ``out=(scores, ids)`` (pinned torch CPU tensors) when given.
        Mirrors what ``KnowledgeBase.search_batch`` hands to ``Dataset.search_batch`` (meerqat/ir/search.py:146)."""
        import numpy as np
        import torch

        from . import _lib

        q = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        if q.dim() != 2:
            raise ValueError("Shape of query must be 2D")
        nq, d = q.shape
        dev = torch.device(f"cuda:{self.local_index.device}")
        G, r = self.world_size, self.rank
        tdt = self._transport_dtype()
        l2 = bool(getattr(self.local_index, "do_L2norm", False))
        sliced = False
        if G == 1 or nq < G or self.dist.get_backend(self.group) != "nccl":
            q_all = q.to(dev, non_blocking=True)
        else:
            per = (nq + G - 1) // G
            a, b = min(nq, r * per), min(nq, (r + 1) * per)
            q_all = torch.empty((G * per, d), dtype=tdt, device=dev)
            part = torch.zeros((per, d), dtype=tdt, device=dev)  # the last rank's slice may be short: zero rows
            if b > a:
                mine = q[a:b].to(dev, non_blocking=True)
                with torch.cuda.device(dev):
                    code = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
                    _lib.check(_lib.load().vq_prep_rows(mine.data_ptr(), code[mine.dtype], part.data_ptr(), code[tdt], b - a, d, int(l2),
                                                        torch.cuda.current_stream(dev).cuda_stream))
            xq = self._exchange(queries_window=True) if (part.numel() * part.element_size()) % 16 == 0 else None
            if xq is not None and part.numel() * part.element_size() <= self.PEER_SLOT_BYTES:
                with torch.cuda.device(dev):
                    stream = torch.cuda.current_stream(dev).cuda_stream
                    xq.publish(part, stream)
                    xq.gather(part.numel() * part.element_size(), q_all, stream)
            else:
                self.dist.all_gather_into_tensor(q_all, part, group=self.group)
            q_all = q_all[:nq]
            sliced = True
        # a slice was normalised in fp32 BEFORE it was rounded for transport (what a single GPU does): not again
        s, i = self.search_tensors(q_all, k, l2norm=False if (sliced and l2) else None)
        if result_ranks != "all" and int(result_ranks) != r:
            return None
        if out is not None:
            out[0].copy_(s, non_blocking=True)
            out[1].copy_(i, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out
        return s.cpu(), i.cpu()

    def search_batch(self, queries, k: int = 100):
        """numpy/torch host queries [nq, d] -> numpy (scores, ids) on every rank, like ``KnowledgeBase.search_batch``
        expects from ``Dataset.search_batch`` (meerqat/ir/search.py:146)."""
        s, i = self.search_host(queries, k)
        return s.numpy(), i.numpy()


class CapturedShardedSearch:
    """``ShardedFlatIndex.capture_search``: write ``queries``, ``replay()``, read ``scores`` / ``ids`` (static device
    tensors).  Every rank must replay the same number of times (the exchange is collective)."""

    def __init__(self, sharded: ShardedFlatIndex, nq: int, k: int, query_dtype=None):
        import torch

        from . import _lib
        from .index import _tensor_dtype_code, packed_result_buffer
        idx = sharded.local_index
        xchg = sharded._exchange()
        if xchg is None or k > 128:
            raise RuntimeError("capture_search needs the NVLink peer-memory exchange (one node, k <= 128)")
        self.sharded, self.nq, self.k = sharded, nq, k
        dev = torch.device(f"cuda:{idx.device}")
        self.queries = torch.zeros((nq, idx.d), dtype=query_dtype or torch.float32, device=dev)
        self._mine, s_v, i_v = packed_result_buffer(nq, k, dev)
        self.scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        self.ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        lib = _lib.load()
        self._ws = torch.empty((lib.vq_index_workspace_bytes(idx._h, nq, k),), dtype=torch.uint8, device=dev)
        l2 = int(idx.do_L2norm)
        qcode = _tensor_dtype_code(self.queries)

        def launch(stream):
            _lib.check(lib.vq_index_search(idx._h, self.queries.data_ptr(), nq, qcode, l2, k, s_v[0].data_ptr(), i_v[0].data_ptr(),
                                           self._ws.data_ptr(), self._ws.numel(), stream.cuda_stream))
            xchg.publish(self._mine, stream.cuda_stream)
            xchg.merge_topk(nq, k, k, self.scores, self.ids, stream.cuda_stream)

        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch(side)  # warm-up outside capture (one real exchange on every rank)
            torch.cuda.current_stream(dev).wait_stream(side)
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                launch(torch.cuda.current_stream(dev))

    def replay(self):
        self.graph.replay()
        return self.scores, self.ids

    def __call__(self, queries):
        self.queries.copy_(queries, non_blocking=True)
        return self.replay()


class ShardedB200Index:
    """What ``KnowledgeBase`` registers for ``"kind": "B200"`` with ``"device": -1`` / ``[g0, g1, ...]``: this rank's
    row shard + the exchange, behind the same surface as ``B200FlatIndex`` (``search_batch`` -> host results on
    EVERY rank, ``search_tensors``, ``do_L2norm``, ``device``, ``len`` = rows of the whole corpus), so that
    ``KnowledgeBase.search_batch`` / ``Searcher`` / ``Dataset.search_batch`` run unchanged under ``torchrun``.
    Every rank runs the same program on the same query batches (the reference's ``Dataset.map`` loop, executed
    once per process); the collectives inside keep the ranks in step."""

    def __init__(self, local_index, group=None):
        import torch
        import torch.distributed as dist

        self.local_index = local_index
        self.sharded = ShardedFlatIndex(local_index, group=group)
        self.device = local_index.device
        self.d = local_index.d
        self.dtype = local_index.dtype
        n = torch.tensor([local_index.ntotal], dtype=torch.int64, device=f"cuda:{self.device}")
        dist.all_reduce(n, group=group)
        self._ntotal = int(n.item())

    @property
    def do_L2norm(self):
        return self.local_index.do_L2norm

    @do_L2norm.setter
    def do_L2norm(self, v):
        self.local_index.do_L2norm = bool(v)

    @property
    def ntotal(self):
        return self._ntotal

    def __len__(self):
        return self._ntotal

    def __deepcopy__(self, memo):  # HF datasets deep-copies Dataset._indexes: the index is a handle on device memory
        return self

    def __copy__(self):
        return self

    def search_tensors(self, queries, k: int = 100, l2norm=None, out=None):
        s, i = self.sharded.search_tensors(queries, k, l2norm=l2norm)
        if out is not None:
            out[0].copy_(s)
            out[1].copy_(i)
            return out
        return s, i

    def search_batch(self, queries, k: int = 10, **kwargs):
        import numpy as np
        import torch

        from .index import BatchedSearchResults
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            return BatchedSearchResults(*self.search_tensors(queries, k))
        queries = np.asarray(queries, dtype=np.float32)
        if len(queries.shape) != 2:
            raise ValueError("Shape of query must be 2D")  # datasets/search.py:380-381
        if queries.shape[1] != self.d:
            raise ValueError(f"dimension mismatch: index d={self.d}, queries d={queries.shape[1]}")
        s, i = self.sharded.search_host(np.ascontiguousarray(queries), k)
        return BatchedSearchResults(s.numpy(), i.numpy().astype(int))

    def search(self, query, k: int = 10, **kwargs):
        import numpy as np

        from .index import SearchResults
        query = np.asarray(query, dtype=np.float32)
        if len(query.shape) != 1 and (len(query.shape) != 2 or query.shape[0] != 1):
            raise ValueError("Shape of query is incorrect, it has to be either a 1D array or 2D (1, N)")
        res = self.search_batch(query.reshape(1, -1), k)
        return SearchResults(res.total_scores[0], res.total_indices[0])

    def shard_file(self, file) -> str:
        return f"{file}.shard{self.sharded.rank}of{self.sharded.world_size}"

    def save(self, file, storage_options=None):
        """Every rank writes its own shard next to ``file`` (``<file>.shard<r>of<G>`` + ``.npy``)."""
        self.local_index.save(self.shard_file(file))

    def close(self):
        self.local_index.close()
