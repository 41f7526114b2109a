"""The multi-rank exchange step (row shards -> all-gather of per-rank candidates -> k-way merge) on
CPU with the gloo backend, world_size 2.  The CUDA engine is replaced by the oracle through the
injection points of ShardedFlatIndex (test-only); the collective plumbing is the product code."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import flat_ip
    from viquae_b200.sharded import ShardedFlatIndex, shard_range

    rng = np.random.default_rng(11)
    X = rng.standard_normal((1001, 24), dtype=np.float32)
    Q = rng.standard_normal((5, 24), dtype=np.float32)
    a, b = shard_range(len(X), rank, world)

    def local_search(q, k):
        D, I = flat_ip.flat_ip_search(q.numpy(), X[a:b], k, id_offset=a)
        return torch.from_numpy(D), torch.from_numpy(I)

    def merge(s_all, i_all, k):
        D, I = flat_ip.merge_topk(list(s_all.numpy()), list(i_all.numpy()), k)
        return torch.from_numpy(D), torch.from_numpy(I)

    idx = ShardedFlatIndex(None, local_search=local_search, merge=merge)
    s, i = idx.search_tensors(torch.from_numpy(Q), 15)
    D, I = flat_ip.flat_ip_search(Q, X, 15)
    ok = np.array_equal(i.numpy(), I) and np.array_equal(s.numpy(), D)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def _worker_gather(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viquae_b200.inbatch import gather_global_batch

    N, M, d = 3, 2, 4
    q = torch.full((N, d), float(rank))
    c = torch.full((N * M, d), float(10 + rank))
    labels = torch.tensor([0, -100, 4])
    gq, gc, gl = gather_global_batch(q, c, labels)
    # trainee.py:446-452: labels of rank i are shifted by i * N*M, ignored labels stay
    want = torch.tensor([0, -100, 4, 0 + N * M, -100, 4 + N * M])
    ok = gq.shape == (world * N, d) and gc.shape == (world * N * M, d) and torch.equal(gl, want)
    ok = ok and bool((gq[N:] == 1).all()) and bool((gc[:N * M] == 10).all()) and bool((gc[N * M:] == 11).all())
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_global_batch_gather_shifts_labels_world2_gloo():
    """The all-gather + label shift of the bi-encoder's in-batch step (meerqat/train/trainee.py:438-452), world size 2."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_gather, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
