"""Rows next to the search path: the bi-encoder's in-batch similarity (forward / evaluation step) against
a plain PyTorch fp32 restatement of meerqat/train/trainee.py:454-458 + meerqat/train/metrics.py:22-37, and hard-negative
mining (EXPERIMENTS.rst:78-103) against set arithmetic on the oracle's results."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def reference_step(q, c, labels, ignore_index=-100):
    """What BiEncoder.step + batch_retrieval compute (torch float32, the reference's own operations)."""
    sims = q.float() @ c.float().T
    log_probs = torch.log_softmax(sims, dim=1)
    loss = torch.nn.functional.nll_loss(log_probs, labels, ignore_index=ignore_index)
    mrr, hits, ignored = 0.0, 0, 0
    rankings = (-log_probs).argsort(dim=1, stable=True)
    for ranking, label in zip(rankings, labels):
        if label == ignore_index:
            ignored += 1
            continue
        if ranking[0] == label:
            hits += 1
        mrr += 1.0 / ((ranking == label).nonzero()[0].item() + 1)
    return log_probs, float(loss), {"MRR@N*M": mrr, "hits@1": hits, "ignored_predictions": ignored, "batch_size": q.shape[0]}


@pytest.mark.parametrize("dtype,N,M,d", [(torch.float32, 32, 2, 768), (torch.float32, 256, 2, 768), (torch.bfloat16, 64, 4, 512), (torch.float16, 7, 3, 100)])
def test_inbatch_step_equals_torch_restatement(dtype, N, M, d):
    from viquae_b200.inbatch import inbatch_step
    g = torch.Generator(device="cuda").manual_seed(3)
    q = (torch.randn((N, d), generator=g, device="cuda") * 0.2).to(dtype)
    c = (torch.randn((N * M, d), generator=g, device="cuda") * 0.2).to(dtype)
    c[::M] += q * 2.0  # the positive of question i sits at i*M, like the reference's batches
    labels = torch.arange(N, device="cuda") * M
    labels[1] = -100
    if N > 5:
        labels[5] = -100
    out = inbatch_step(q, c, labels)
    ref_lp, ref_loss, ref_metrics = reference_step(q, c, labels)
    # fp32 accumulation in a different order than cuBLAS: 1e-5 relative on the similarities (tolerance of north_star's fp32 mode)
    assert torch.allclose(out["log_probs"], ref_lp, rtol=1e-5, atol=2e-5)
    assert abs(out["loss"] - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert out["metrics"]["hits@1"] == ref_metrics["hits@1"] and out["metrics"]["ignored_predictions"] == ref_metrics["ignored_predictions"]
    assert out["metrics"]["batch_size"] == N and abs(out["metrics"]["MRR@N*M"] - ref_metrics["MRR@N*M"]) < 1e-9
    assert (out["ranks"][labels == -100] == 0).all()


def test_mine_hard_negatives_splits_the_retrieved_ids():
    from viquae_b200.index import B200FlatIndex
    from viquae_b200.inbatch import mine_hard_negatives
    rng = np.random.default_rng(13)
    X = rng.standard_normal((5000, 64)).astype(np.float32)
    Q = rng.standard_normal((300, 64)).astype(np.float32)
    k = 20
    idx = B200FlatIndex(d=64, dtype="f32")
    idx.add(X)
    D, I = c_oracle.flat_ip(Q, X, k)
    relevant = [sorted(set(rng.choice(5000, size=30, replace=False).tolist()) | set(I[i, 3:9:2].tolist())) if i % 3 else [] for i in range(len(Q))]
    fallback = [[-7, i] for i in range(len(Q))]
    prov, irr = mine_hard_negatives(idx, torch.from_numpy(Q).cuda(), relevant, k=k, fallback=fallback, batch_size=128)
    for i in range(len(Q)):
        rel = set(relevant[i])
        want_rel = [int(x) for x in I[i] if x in rel]
        assert prov[i] == (want_rel if want_rel else fallback[i]), i
        assert irr[i] == [int(x) for x in I[i] if x not in rel], i   # hard negatives, best first
    idx.close()
