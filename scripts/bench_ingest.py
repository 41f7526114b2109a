"""Ingest rate: host fp32 embeddings -> bf16 index rows on the device, through
B200FlatIndex.add (numpy blocks), add_vectors over an Arrow-backed HF Dataset column (zero-copy blocks),
and device-resident tensors.  Prints one JSON line per path.
python scripts/bench_ingest.py [--rows 2000000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viquae_b200.index import B200FlatIndex

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000)
ap.add_argument("--arrow-rows", type=int, default=400_000)
ap.add_argument("--d", type=int, default=768)
args = ap.parse_args()
d = args.d
rng = np.random.default_rng(0)
block = rng.standard_normal((250_000, d), dtype=np.float32)
x = np.concatenate([block] * (args.rows // 250_000), axis=0)
gb = x.nbytes / 1e9


def timed(label, fn, nbytes, rows):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert len(idx) == rows
    print(json.dumps({"path": label, "rows": rows, "source_GB": round(nbytes / 1e9, 3), "seconds": round(dt, 3),
                      "source_GB_per_s": round(nbytes / 1e9 / dt, 2), "rows_per_s": round(rows / dt)}), flush=True)
    idx.close()


def numpy_blocks(l2norm):
    def fn():
        idx = B200FlatIndex(d=d, dtype="bf16", string_factory="L2norm,Flat" if l2norm else "Flat", capacity=len(x))
        for i in range(0, len(x), 262_144):
            idx.add(x[i:i + 262_144])
        return idx
    return fn


def one_call():
    idx = B200FlatIndex(d=d, dtype="bf16", capacity=len(x))
    idx.add(x)
    return idx


def device_tensor():
    idx = B200FlatIndex(d=d, dtype="bf16", capacity=len(x))
    for i in range(0, len(x), 250_000):
        idx.add(xd)
    return idx


numpy_blocks(False)().close()  # warm-up (context, module load)
timed("numpy fp32 host, 262144-row blocks -> bf16", numpy_blocks(False), x.nbytes, len(x))
timed("numpy fp32 host, 262144-row blocks -> L2norm + bf16", numpy_blocks(True), x.nbytes, len(x))
timed("numpy fp32 host, one call -> bf16", one_call, x.nbytes, len(x))
xd = torch.from_numpy(block).cuda()
timed("device fp32 tensor blocks -> bf16", device_tensor, x.nbytes, len(x))
try:
    import datasets
    n = args.arrow_rows
    ds = datasets.Dataset.from_dict({"emb": [row for row in x[:n]]}) if False else None
    import pyarrow as pa
    flat = pa.array(x[:n].reshape(-1))
    col = pa.FixedSizeListArray.from_arrays(flat, d)
    ds = datasets.Dataset(pa.table({"emb": col}))

    def arrow():
        idx = B200FlatIndex(d=d, dtype="bf16", capacity=n)
        idx.add_vectors(ds, column="emb")
        return idx
    timed("HF Dataset Arrow column (add_vectors, zero-copy blocks) -> bf16", arrow, x[:n].nbytes, n)
except Exception as e:  # datasets/pyarrow layout differences are not the engine's concern
    print(json.dumps({"path": "arrow", "skipped": repr(e)}))
