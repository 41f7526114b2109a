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
