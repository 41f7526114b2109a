"""Per-call latency of small-batch searches with and without CUDA-graph replay (BASELINE config 5).
python scripts/graph_latency.py [--rows 1500000] [--calls 300]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viquae_b200.index import B200FlatIndex

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="+", default=[1_500_000, 12_000_000])
ap.add_argument("--calls", type=int, default=1000)
ap.add_argument("--warmup", type=int, default=50)
ap.add_argument("--nq", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32])
args = ap.parse_args()
dev = torch.device("cuda:0")
for rows in args.rows:
    idx = B200FlatIndex(d=768, dtype="bf16", capacity=rows)
    for c0 in range(0, rows, 1_000_000):
        g = torch.Generator(device=dev); g.manual_seed(20240117 + c0 // 1_000_000)
        idx.add(torch.randn((min(1_000_000, rows - c0), 768), generator=g, device=dev).to(torch.bfloat16))
    for nq in args.nq:
        g = torch.Generator(device=dev); g.manual_seed(7)
        q = torch.randn((nq, 768), generator=g, device=dev)
        out = (torch.empty((nq, 100), device=dev), torch.empty((nq, 100), dtype=torch.int64, device=dev))
        cap = idx.capture_search(nq, 100)
        cap.queries.copy_(q)

        def timed(fn):
            for _ in range(args.warmup):
                fn()
            torch.cuda.synchronize()
            lat = []
            for _ in range(args.calls):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                lat.append(e0.elapsed_time(e1) * 1e3)
            lat.sort()
            return lat[len(lat) // 2], lat[int(len(lat) * 0.99)]
        p_plain = timed(lambda: idx.search_tensors(q, 100, out=out))
        p_graph = timed(cap.replay)
        same = bool(torch.equal(out[1], cap.ids) and torch.equal(out[0], cap.scores))
        ideal = rows * 768 * 2 / 6540.5e9 * 1e6
        print(json.dumps({"rows": rows, "nq": nq, "plain_p50_us": round(p_plain[0], 1), "plain_p99_us": round(p_plain[1], 1),
                          "graph_p50_us": round(p_graph[0], 1), "graph_p99_us": round(p_graph[1], 1),
                          "graph_queries_per_s": round(nq / p_graph[0] * 1e6), "calls": args.calls, "warmup": args.warmup,
                          "hbm_ideal_us": round(ideal, 1), "identical": same,
                          "ids_sum": int(cap.ids.sum()), "scores_sum": float(cap.scores.double().sum())}), flush=True)
    idx.close()
