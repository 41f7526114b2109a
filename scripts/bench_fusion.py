"""BASELINE configs[2] end to end on one GPU: DPR text (768-d, 12M passages... here --text-rows) + CLIP image
(512-d x 1,495,352, "L2norm,Flat") + ArcFace face (512-d x 506,237, "L2norm,Flat", ~half of the queries None),
1,250 queries, top-100 per index, article->passage expansion of the image/face hits (mean 8 passages per
article), then default-minimum + gzmuv + weighted sum.  Device-resident pipeline: K0/K3/K4 -> K6 -> K5.
Prints one JSON line with per-stage CUDA-event times.  Synthetic data generated on the device."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viquae_b200.index import B200FlatIndex
from viquae_b200.postprocess import CSRMapping, expand_hits
from viquae_b200.fuse import fuse_dense

ap = argparse.ArgumentParser()
ap.add_argument("--text-rows", type=int, default=12_000_000)
ap.add_argument("--nq", type=int, default=1250)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
k, nq = 100, args.nq


def build(rows, d, factory, seed):
    idx = B200FlatIndex(d=d, dtype="bf16", string_factory=factory, capacity=rows)
    for c0 in range(0, rows, 1_000_000):
        g = torch.Generator(device=dev); g.manual_seed(seed + c0)
        idx.add(torch.randn((min(1_000_000, rows - c0), d), generator=g, device=dev).to(torch.bfloat16))
    return idx

text = build(args.text_rows, 768, "Flat", 1)
image = build(1_495_352, 512, "L2norm,Flat", 2)
face = build(506_237, 512, "L2norm,Flat", 3)
g = torch.Generator(device=dev); g.manual_seed(7)
q_text = torch.randn((nq, 768), generator=g, device=dev)
q_img = torch.randn((nq, 512), generator=g, device=dev)
has_face = torch.rand((nq,), generator=g, device=dev) < 0.5
q_face = torch.randn((int(has_face.sum()), 512), generator=g, device=dev)
face_rows = torch.nonzero(has_face).flatten()
# article -> passages CSR (8 passages per article on average); passage ids live in the text KB's id space
rng = np.random.default_rng(0)
deg = rng.integers(1, 16, size=1_495_352)
off = np.zeros(len(deg) + 1, np.int64); np.cumsum(deg, out=off[1:])
m = CSRMapping.__new__(CSRMapping)
m.n_rows, m.max_degree = len(deg), int(deg.max())
m.offsets = torch.from_numpy(off).to(dev)
m.targets = (torch.arange(int(off[-1]), device=dev, dtype=torch.int64) % args.text_rows)
weights = [0.4, 0.3, 0.3]


def step(ev):
    ev[0].record()
    s_t, i_t = text.search_tensors(q_text, k)
    s_i, i_i = image.search_tensors(q_img, k)
    s_f, i_f = face.search_tensors(q_face, k)
    ev[1].record()
    r_t = expand_hits(s_t, i_t, None, None, k)
    r_i = expand_hits(s_i, i_i, m, None, k)
    r_f = expand_hits(s_f, i_f, m, None, k)
    cap = max(r_t[0].shape[1], r_i[0].shape[1], r_f[0].shape[1])
    S = torch.zeros((3, nq, cap), dtype=torch.float64, device=dev)
    I = torch.full((3, nq, cap), -1, dtype=torch.int64, device=dev)
    S[0, :, :r_t[0].shape[1]], I[0, :, :r_t[1].shape[1]] = r_t[0], r_t[1]
    S[1, :, :r_i[0].shape[1]], I[1, :, :r_i[1].shape[1]] = r_i[0], r_i[1]
    S[2, face_rows, :r_f[0].shape[1]], I[2, face_rows, :r_f[1].shape[1]] = r_f[0], r_f[1]
    ev[2].record()
    out = fuse_dense(S, I, weights, norm="gzmuv", defmin=True)
    ev[3].record()
    return out

evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps + 3)]
for e in evs[:3]:
    step(e)
torch.cuda.synchronize()
for e in evs[3:]:
    out = step(e)
torch.cuda.synchronize()
t = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs[3:]]).mean(0)
print(json.dumps({"workload": f"config 3: text {args.text_rows}x768 + image 1495352x512 + face 506237x512 (50% None), {nq} queries, top-100, gzmuv+defmin+wsum",
                  "search_ms": t[0], "expand_ms": t[1], "fuse_ms": t[2], "total_ms": float(t.sum()), "queries_per_s": nq / (t.sum() / 1e3),
                  "fused_union_mean": float(out[2].float().mean().item())}))
