"""One 1.5M x 768 bf16 shard, a few nq=1 searches: target for an ncu launch list of the latency regime."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viquae_b200.index import B200FlatIndex
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
idx = B200FlatIndex(d=768, dtype="bf16", capacity=rows)
for c0 in range(0, rows, 500_000):
    idx.add(torch.randn((min(500_000, rows - c0), 768), device=dev).to(torch.bfloat16))
q = torch.randn((nq, 768), device=dev)
for _ in range(4):
    idx.search_tensors(q, 100)
torch.cuda.synchronize()
