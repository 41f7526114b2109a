"""K5 (score fusion) against the float64 oracle and the reference-generated golden fixture."""
import json
import os

import numpy as np
import pytest

from oracle import fuse_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9  # float64 arithmetic, different summation order than numpy's pairwise sums


def assert_runs_close(mine, ref, tol=TOL):
    assert mine.keys() == ref.keys()
    for q in ref:
        assert set(mine[q]) == set(ref[q]), q
        for d, v in ref[q].items():
            assert abs(mine[q][d] - v) <= tol * max(1.0, abs(v)), (q, d, mine[q][d], v)
        # same ranking up to ties within tolerance
        mv = list(mine[q].values())
        assert all(a >= b - 1e-12 for a, b in zip(mv, mv[1:]))


def test_reference_golden_defmin_and_gzmuv(golden_dir):
    from viquae_b200 import fuse
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    filled = fuse.default_minimum([fuse.Run(r) for r in g["runs"]])
    for mine, ref in zip(filled, g["after_defmin"]):
        assert_runs_close(mine.run, ref, tol=0.0)   # fill values are copies: exact
    for run, ref in zip(g["after_defmin"], g["gzmuv_after_defmin"]):
        assert_runs_close(fuse.gzmuv_norm(fuse.Run(run)).run, ref)
    for run, ref in zip(g["runs"], g["gzmuv_raw"]):
        assert_runs_close(fuse.gzmuv_norm(fuse.Run(run)).run, ref)


@pytest.mark.parametrize("norm", ["zmuv", "gzmuv", "min-max", None])
@pytest.mark.parametrize("defmin", [False, True])
def test_fusion_test_matches_oracle(golden_dir, norm, defmin):
    from viquae_b200 import fuse
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    weights = [0.6, 0.2, 0.2]   # experiments/ir/viquae/**/config_test.json style weights
    fused = fuse.Fusion(runs=[fuse.Run(r) for r in g["runs"]], norm=norm, defmin=defmin).test({"weights": weights}, metrics=None)
    ref = fuse_oracle.fuse(g["runs"], weights, norm=norm, defmin=defmin)
    assert_runs_close(fused.run, ref)


def test_fusion_config3_shape_dense():
    """BASELINE configs[2]: text + image + face runs of 1,250 queries x top-100, ~half of the face queries None."""
    import torch
    from viquae_b200 import fuse
    rng = np.random.default_rng(0)
    R, nq, k = 3, 1250, 100
    scores = np.sort(rng.normal([[[71.3]], [[0.3]], [[0.5]]], [[[2.2]], [[0.1]], [[0.2]]], size=(R, nq, k)), axis=2)[:, :, ::-1].copy()
    ids = np.stack([np.stack([rng.choice(4000, size=k, replace=False) for _ in range(nq)]) for _ in range(R)]).astype(np.int64)
    ids[2, rng.random(nq) < 0.5] = -1
    w = [0.5, 0.3, 0.2]
    out_s, out_i, out_c = fuse.fuse_dense(scores, ids, w, norm="gzmuv", defmin=True)
    ref = fuse_oracle.fuse(fuse_oracle.dense_to_runs(scores, ids), w, norm="gzmuv", defmin=True)
    out_s, out_i, out_c = out_s.cpu().numpy(), out_i.cpu().numpy(), out_c.cpu().numpy()
    for q in range(0, nq, 37):
        n = out_c[q]
        mine = {int(d): float(s) for d, s in zip(out_i[q, :n], out_s[q, :n])}
        assert_runs_close({q: mine}, {q: ref[q]})
    assert (out_i[np.arange(nq)[:, None], np.arange(R * k)[None, :]][out_c[:, None] <= np.arange(R * k)[None, :]] == -1).all()


def test_fit_finds_the_informative_run():
    from viquae_b200 import fuse
    good = {f"q{i}": {"1": 1.0, "2": 0.5, "3": 0.1} for i in range(6)}
    bad = {f"q{i}": {"3": 1.0, "2": 0.5, "1": 0.1} for i in range(6)}
    qrels = fuse.Qrels({f"q{i}": {"1": 1} for i in range(6)})
    res = fuse.Fusion(qrels=qrels, runs=[fuse.Run(bad), fuse.Run(good)], norm="zmuv").fit("mrr@100")
    (params, score), = res.values()
    assert score == 1.0 and params["weights"][1] > params["weights"][0]


def test_l2norm_host_api():
    from viquae_b200.search import L2norm
    x = np.array([[3, 4, 0], [0, 0, 0]], np.float32)
    out = L2norm(x)
    assert np.allclose(out[0], [0.6, 0.8, 0]) and np.isnan(out[1]).all()


def test_fit_dense_metric_equals_host_metric(golden_dir):
    """The on-device metric used by Fusion.fit's grid search equals the dict-based evaluate() on the same fused run."""
    from viquae_b200 import fuse, runs
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    rr = [fuse.Run(r) for r in g["runs"]]
    qrels = runs.Qrels({q: {d: 1 for d in list(res)[2:4]} for q, res in g["runs"][0].items()})
    f = fuse.Fusion(qrels=qrels, runs=rr, norm="zmuv", defmin=True)
    res = f.fit("mrr@100")
    (params, score), = res.values()
    host = runs.evaluate(qrels, f._fused(params["weights"], "zmuv"), "mrr@100")
    assert abs(host - score) < 1e-12
    best_host = max(runs.evaluate(qrels, f._fused(w, "zmuv"), "mrr@100") for w in fuse.weight_grid(3))
    assert abs(best_host - score) < 1e-12
