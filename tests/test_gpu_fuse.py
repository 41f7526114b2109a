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


@pytest.mark.parametrize("norm", ["zmuv", "gzmuv", "min-max", None])
def test_grid_search_kernel_equals_host_evaluation_at_every_grid_point(golden_dir, norm):
    """vq_fuse_grid (ranks of the relevant documents by counting, no sort) vs the dict-based evaluate() of the run
    that vq_fuse produces for the same weights — for every point of the 66-point grid and every metric the reference
    reports (search.py:397), with a judged query that no run answers and a relevant document nobody retrieved."""
    from viquae_b200 import fuse, runs
    g = json.load(open(os.path.join(golden_dir, "fusion_golden.json")))
    rr = [fuse.Run(r) for r in g["runs"]]
    q = {qid: {d: 1 for d in list(res)[1:6:2]} for qid, res in g["runs"][0].items()}
    q["q1"]["never-retrieved"] = 1
    q["q2"] = {d: 0 for d in q["q2"]}            # judged, nothing relevant
    q["not-in-any-run"] = {"5": 1}
    qrels = runs.Qrels(q)
    f = fuse.Fusion(qrels=qrels, runs=rr, norm=norm, defmin=True)
    for metric in ("mrr@100", "precision@1", "precision@5", "hit_rate@3", "recall@10", "mrr@2"):
        res = f.fit(metric)
        (params, score), = res.values()
        report = f.grid_reports[(norm, "wsum")]
        assert len(report) == 66
        host = [runs.evaluate(qrels, f._fused(w, norm), metric) for w, _ in report]
        dev = [v for _, v in report]
        assert np.allclose(dev, host, rtol=0, atol=1e-12), (metric, np.abs(np.array(dev) - np.array(host)).max())
        first_best = int(np.argmax(np.array(host) > max(host) - 1e-15))
        assert params["weights"] == [round(w, 10) for w in report[first_best][0]] and abs(score - host[first_best]) < 1e-12


def test_fusion_beyond_the_old_2048_entry_limit():
    """n_runs * k_in above 2048 (a hit on a long article expands to hundreds of passages): served from dynamic
    shared memory up to 16384 entries, equal to the oracle."""
    from oracle import fuse_oracle
    from viquae_b200 import fuse
    rng = np.random.default_rng(8)
    runs_ = []
    for r in range(3):
        run = {}
        for qi in range(5):
            docs = rng.choice(4000, size=900, replace=False)
            run[f"q{qi}"] = {str(int(d)): float(s) for d, s in zip(docs, np.sort(rng.normal(0, 1, 900))[::-1])}
        runs_.append(run)
    fused = fuse.Fusion(qrels=None, runs=[fuse.Run(r) for r in runs_], norm="zmuv", defmin=True).test({"weights": [0.5, 0.3, 0.2]}, metrics=[])
    ref = fuse_oracle.fuse(runs_, [0.5, 0.3, 0.2], norm="zmuv", defmin=True)
    for qid in ref:
        assert set(fused.run[qid]) == set(ref[qid])
        assert np.allclose([fused.run[qid][d] for d in ref[qid]], list(ref[qid].values()), rtol=1e-9, atol=1e-9)
