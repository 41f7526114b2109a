"""Row-sharded search over 2 GPUs with NCCL (skipped on a 1-GPU box): equals the single-index result."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["VQ_ROOT"])
import numpy as np, torch, torch.distributed as dist
from viquae_b200.index import B200FlatIndex
from viquae_b200.sharded import ShardedFlatIndex, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
rng = np.random.default_rng(5)
X = rng.standard_normal((30011, 256), dtype=np.float32); Q = rng.standard_normal((77, 256), dtype=np.float32)
a, b = shard_range(len(X), rank, world)
local = B200FlatIndex(d=256, dtype="bf16", device=rank, id_offset=a); local.add(X[a:b])
s, i = ShardedFlatIndex(local).search_tensors(torch.from_numpy(Q).cuda(), 100)
if rank == 0:
    full = B200FlatIndex(d=256, dtype="bf16", device=0); full.add(X)
    s0, i0 = full.search_tensors(torch.from_numpy(Q).cuda(), 100)
    assert torch.equal(i, i0) and torch.allclose(s, s0, rtol=1e-5), "sharded != single"
sh = ShardedFlatIndex(local)
sh.SLICE_EXCHANGE_MIN_QUERIES = 16          # force the all-to-all-by-query-slice exchange (64k-batch path)
s2, i2 = sh.search_tensors(torch.from_numpy(Q).cuda(), 100)
assert torch.equal(i2, i) and torch.equal(s2, s), "slice exchange != all-gather exchange"
# transports: NVLink peer-memory publish + pull-merge (default) vs NCCL all-gather + merge kernel
peer, nccl = ShardedFlatIndex(local, transport="peer"), ShardedFlatIndex(local, transport="nccl")
Qd = torch.from_numpy(Q).cuda()
for rep in range(5):                        # several sequence numbers: both slot parities, flags only grow
    qs = Qd.roll(rep, 0)[: 77 - 9 * rep].contiguous()
    sp, ip = peer.search_tensors(qs, 100)
    sn, i_n = nccl.search_tensors(qs, 100)
    assert peer._xchg is not None and nccl._xchg is None
    assert torch.equal(ip, i_n) and torch.equal(sp, sn), f"peer transport != nccl transport (rep {rep})"
assert int((ip < 0).sum()) == 0
# host entry point: every rank uploads a slice of the queries, the slices travel through the second window
sh_h, ih_h = peer.search_host(Q, 100)
s_f, i_f = peer.search_tensors(Qd, 100)
assert torch.equal(ih_h, i_f.cpu()) and torch.equal(sh_h, s_f.cpu()), "search_host (sliced upload) != search_tensors"
only0 = peer.search_host(Q, 100, result_ranks=0)
assert (only0 is None) == (rank != 0)
# the whole sharded step as one CUDA graph per rank
cap = peer.capture_search(13, 100, query_dtype=torch.float32)
for rep in range(4):
    cap.queries.copy_(Qd[rep:rep + 13])
    sc, ic = cap.replay()
    se, ie = peer.search_tensors(Qd[rep:rep + 13].contiguous(), 100)
    assert torch.equal(ic, ie) and torch.equal(sc, se), f"graph replay {rep} != direct"
if rank == 0:
    print("MULTI_OK")
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpu_sharded_search(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(SCRIPT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, VQ_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert "MULTI_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


CONFIG_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, os.environ["VQ_ROOT"])
import numpy as np, torch, torch.distributed as dist
import datasets
from viquae_b200.search import KnowledgeBase, dataset_search
from viquae_b200.sharded import ShardedB200Index
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
tmp = os.environ["VQ_TMP"]
rng = np.random.default_rng(31)
n, d1, d2, nq, k = 20011, 64, 32, 70, 20
txt = rng.standard_normal((n, d1)).astype(np.float32)
img = rng.standard_normal((n, d2)).astype(np.float32)
if rank == 0:
    datasets.Dataset.from_dict({"text_emb": txt.tolist(), "image_emb": img.tolist()}).save_to_disk(os.path.join(tmp, "kb"))
    json.dump({f"q{j}": {"0": 1} for j in range(nq)}, open(os.path.join(tmp, "qrels.json"), "wt"))
dist.barrier()
q1 = rng.standard_normal((nq, d1)).astype(np.float32); q2 = rng.standard_normal((nq, d2)).astype(np.float32)
queries = datasets.Dataset.from_dict({"id": [f"q{j}" for j in range(nq)], "q_text": q1.tolist(), "q_image": q2.tolist()})

def config(dev_a, dev_b):
    return {os.path.join(tmp, "kb"): {"index_kwargs": {
        "dpr": {"column": "text_emb", "kind": "B200", "key": "q_text", "string_factory": "Flat", "metric_type": 0, "dtype": "f32x3", "device": dev_a},
        "clip": {"column": "image_emb", "kind": "B200", "key": "q_image", "string_factory": "L2norm,Flat", "metric_type": 0, "dtype": "bf16", "device": dev_b}}}}

fusion = {"subcommand": "test", "subcommand_kwargs": {"best_params": {"weights": [0.6, 0.4]}}, "norm": "zmuv", "defmin": True}
out = os.path.join(tmp, "metrics_sharded")
sh = dataset_search(queries, k=k, metric_save_path=out, kb_kwargs=config(-1, list(range(world))), qrels=os.path.join(tmp, "qrels.json"),
                    fusion_kwargs=fusion, map_kwargs={"batch_size": 32})
kb = next(iter(sh.kbs.values()))
assert all(isinstance(e, ShardedB200Index) for e in kb._engines.values()) and len(kb._engines["dpr"]) == n
s_sh, i_sh = kb.search_batch("dpr", q1, k=k)          # merged result on EVERY rank
res = kb.dataset.search_batch("clip", q2, k)           # HF entry point -> the sharded engine
# save_path / load under sharding: every rank writes and reads its own shard file
shard_path = os.path.join(tmp, "dpr.index")
kb._engines["dpr"].save(shard_path)
dist.barrier()
assert os.path.exists(f"{shard_path}.shard{rank}of{world}")
cfg_load = config(-1, list(range(world)))
cfg_load[os.path.join(tmp, "kb")]["index_kwargs"] = {"dpr": {"column": "text_emb", "kind": "B200", "key": "q_text", "string_factory": "Flat",
                                                              "metric_type": 0, "device": -1, "load": True, "file": shard_path}}
kb_loaded = KnowledgeBase(os.path.join(tmp, "kb"), **cfg_load[os.path.join(tmp, "kb")])
s_ld, i_ld = kb_loaded.search_batch("dpr", q1, k=k)
assert np.array_equal(i_ld, i_sh) and np.array_equal(s_ld, s_sh), "sharded index loaded from its shard files != the one that wrote them"
if rank == 0:
    one = dataset_search(queries, k=k, metric_save_path=os.path.join(tmp, "metrics_one"), kb_kwargs=config(0, 0),
                         qrels=os.path.join(tmp, "qrels.json"), fusion_kwargs=fusion, map_kwargs={"batch_size": 32})
    kb1 = next(iter(one.kbs.values()))
    s_1, i_1 = kb1.search_batch("dpr", q1, k=k)
    assert np.array_equal(i_sh, i_1) and np.array_equal(s_sh, s_1), "sharded search_batch != single GPU"
    r1 = kb1.dataset.search_batch("clip", q2, k)
    assert np.array_equal(np.asarray(res.total_indices), np.asarray(r1.total_indices)) and np.array_equal(np.asarray(res.total_scores), np.asarray(r1.total_scores))
    for name in ("dpr", "clip"):
        assert sh.runs[name].run == one.runs[name].run, f"run {name} differs"
    assert sh.fusion_result.run == one.fusion_result.run, "fused run differs"
    assert sorted(os.listdir(out)) == sorted(os.listdir(os.path.join(tmp, "metrics_one")))
    print("CONFIG_SHARDED_OK")
else:
    torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
'''


def test_sharding_selected_from_the_config(tmp_path):
    """`"device": -1` / `[0, 1]` with `"kind": "B200"` in an experiment config row-shards the index over the ranks of a
    torchrun launch; `dataset_search` with a two-index config + fusion then equals the single-GPU run, file for file."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "cfg.py"
    script.write_text(CONFIG_SCRIPT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, VQ_ROOT=ROOT, VQ_TMP=str(tmp_path))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert "CONFIG_SHARDED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
