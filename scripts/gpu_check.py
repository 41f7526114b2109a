"""First-light / debugging harness: runs one engine variant against the numpy oracle and prints a
tie-aware diff.  Usage: python scripts/gpu_check.py <variant> [N] [nq] [d] [k] [dtype]
Test infrastructure (imports oracle/)."""
import sys, time
import numpy as np
import torch

sys.path.insert(0, ".")
from viquae_b200.index import B200FlatIndex
from oracle import flat_ip, c_oracle

variant = sys.argv[1] if len(sys.argv) > 1 else "tc_cg1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 70
d = int(sys.argv[4]) if len(sys.argv) > 4 else 768
k = int(sys.argv[5]) if len(sys.argv) > 5 else 100
dtype = sys.argv[6] if len(sys.argv) > 6 else ("f32" if variant == "f32" else "bf16")

rng = np.random.default_rng(1)
X = rng.standard_normal((N, d), dtype=np.float32)
Q = rng.standard_normal((nq, d), dtype=np.float32)
if dtype == "bf16":
    X, Q = flat_ip.bf16_round(X), flat_ip.bf16_round(Q)
elif dtype == "f16":
    X, Q = flat_ip.fp16_round(X), flat_ip.fp16_round(Q)
idx = B200FlatIndex(d=d, dtype=dtype, device=0, variant=None if variant in ("f32", "auto") else variant)
idx.add(X)
t = time.time()
res = idx.search_batch(Q, k)
torch.cuda.synchronize()
print(f"[{variant}] N={N} nq={nq} d={d} k={k} dtype={dtype} search_host {1e3 * (time.time() - t):.1f} ms", flush=True)
D, I = np.asarray(res.total_scores), np.asarray(res.total_indices)
if dtype == "f32":
    Dr, Ir = c_oracle.flat_ip(Q, X, k)
    print("  bit-exact scores:", bool((D.view(np.uint32) == Dr.view(np.uint32)).all()), " ids equal:", bool((I == Ir).all()))
else:
    Dr, Ir = flat_ip.flat_ip_search(Q, X, k)
cmp = flat_ip.compare_results(Dr, Ir, D, I, rel_tol=2e-3 if dtype != "f32" else 1e-5)
print("  ", cmp, flush=True)
if cmp["real_mismatches"] or cmp["overlap"] < 0.999:
    bad = np.flatnonzero((I != Ir).any(axis=1))[:3]
    for r in bad:
        print("  row", r, "got", I[r][:8], D[r][:8], "\n        ref", Ir[r][:8], Dr[r][:8])
    sys.exit(1)
print("  OK")
