"""Host-side logic that needs no GPU: config parsing, sharding arithmetic, run containers, dense
conversion, metrics, and the experiment configs of the reference (when its checkout is present)."""
import glob
import json
import os

import numpy as np
import pytest

from viquae_b200 import fuse, runs, search, sharded
from viquae_b200.index import parse_string_factory


def test_string_factory_semantics():
    assert parse_string_factory("Flat", 0) is False
    assert parse_string_factory("L2norm,Flat", 0) is True
    assert parse_string_factory("L2norm, Flat", None) is True
    assert parse_string_factory(None, 0) is False
    with pytest.raises(ValueError, match="L2"):
        parse_string_factory(None, None)  # datasets would silently build an L2 index
    with pytest.raises(ValueError):
        parse_string_factory("Flat", 1)
    with pytest.raises(ValueError):
        parse_string_factory("IVF4096,Flat", 0)


def test_index_kind_lookup_by_name():
    assert search.IndexKind["B200"] is search.IndexKind.B200
    assert [k.name for k in search.IndexKind][:3] == ["FAISS", "ES", "PYSERINI"]
    kb = search.KnowledgeBase(load_dataset=False)
    kb.add_or_load_index(column=None, index_name="dpr", kind="B200", key="q_text", string_factory="L2norm,Flat")
    assert kb.indexes["dpr"].do_L2norm is True and kb.indexes["dpr"].key == "q_text"
    with pytest.raises(NotImplementedError):
        kb.add_or_load_index(column=None, index_name="bm25", kind="ES", key="input")
    with pytest.raises(KeyError):
        search.IndexKind["ANNOY"]
    with pytest.warns(UserWarning):  # legacy dialect keys are tolerated
        kb.add_or_load_index(column=None, index_name="old", key="q", string_factory="Flat", es=False, kind_str="TEXT", normalization={"mean": 1})
    with pytest.raises(TypeError):
        kb.add_or_load_index(column=None, index_name="bad", key="q", string_factory="Flat", nprobe=3)


def test_shard_ranges_cover_and_balance():
    for n, w in ((12_000_000, 8), (10, 3), (7, 8), (0, 2)):
        r = [sharded.shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_run_orders_descending_and_roundtrips(tmp_path):
    r = runs.Run({"q": {"a": 0.1, "b": 0.9, "c": 0.5}}, name="x")
    assert list(r.run["q"]) == ["b", "c", "a"]
    r.save(tmp_path / "r.json")
    assert runs.Run.from_file(tmp_path / "r.json").run == r.run


def test_metrics():
    q = runs.Qrels({"q1": {"a": 1}, "q2": {"z": 1}})
    r = runs.Run({"q1": {"b": 3.0, "a": 2.0}, "q2": {"x": 1.0}})
    m = runs.evaluate(q, r, ["mrr@100", "precision@1", "hit_rate@20"])
    assert m == {"mrr@100": 0.25, "precision@1": 0.0, "hit_rate@20": 0.5}


def test_weight_grid_sizes():
    assert [len(list(fuse.weight_grid(n))) for n in (2, 3, 4)] == [11, 66, 286]
    assert all(abs(sum(w) - 1) < 1e-9 for w in fuse.weight_grid(3))


def test_dense_runs_roundtrip():
    a = {"q1": {"10": 2.0, "7": 1.0}, "q2": {}}
    b = {"q1": {"7": 5.0}, "q3": {"1": 0.5}}
    d = fuse.DenseRuns([a, b])
    assert d.q_ids == ["q1", "q2", "q3"] and d.scores.shape == (2, 3, 2)
    assert d.ids[0, 0].tolist() == [10, 7] and d.ids[0, 1].tolist() == [-1, -1] and d.ids[1, 2].tolist() == [1, -1]
    assert d.decode(7) == "7"
    d2 = fuse.DenseRuns([{"q": {"doc_a": 1.0, "doc_b": 2.0}}])
    assert d2.decode(int(d2.ids[0, 0, 0])) == "doc_a"


def test_searcher_postprocessing_matches_reference_golden(golden_dir):
    """Searcher._record vs the runs the reference's own Searcher.__call__ produced (search.py:413-440)."""
    g = np.load(os.path.join(golden_dir, "search_golden.npz"))
    meta = json.load(open(os.path.join(golden_dir, "search_golden.json")))

    class KB:
        index_mapping = {int(k): v for k, v in meta["mapping"].items()}
        many2one = None
    s = search.Searcher.__new__(search.Searcher)
    s.k = meta["k"]
    for q in range(g["dpr_ids"].shape[0]):
        run_q = {}
        s._record(KB, run_q, g["dpr_scores"][q], g["dpr_ids"][q])
        ref = meta["runs"]["dpr"][f"q{q}"]
        assert list(run_q) == list(ref)
        assert np.allclose([float(v) for v in run_q.values()], list(ref.values()), rtol=1e-6)
    KB.many2one = "max"
    KB.index_mapping = {0: [5], 1: [5], 2: [6]}
    run_q = {}
    s._record(KB, run_q, [3.0, 2.0, 1.0], [1, 0, 2])
    assert run_q == {"5": 3.0, "6": 1.0}


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_reference_experiment_configs_parse():
    """Every dense index entry of the shipped IR configs is accepted (legacy keys tolerated), SURVEY §4."""
    paths = glob.glob(os.path.join(REFERENCE, "experiments", "ir", "**", "*.json"), recursive=True)
    n_dense = 0
    for path in paths:
        try:
            cfg = json.load(open(path))
        except Exception:
            continue
        if not isinstance(cfg, dict) or "kb_kwargs" not in cfg:
            continue
        for kb_path, kb_kwarg in cfg["kb_kwargs"].items():
            for name, ik in kb_kwarg.get("index_kwargs", {}).items():
                ik = dict(ik)
                if ik.get("es") or ik.get("kind") in ("ES", "PYSERINI"):
                    continue
                ik.pop("column", None)
                kb = search.KnowledgeBase(load_dataset=False)
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    if ik.get("string_factory") is None and ik.get("metric_type") is None:
                        ik["metric_type"] = 0
                    kb.add_or_load_index(column=None, index_name=name, **ik)
                assert name in kb.indexes
                n_dense += 1
    assert n_dense >= 5


def test_arrow_column_blocks_zero_copy_and_fallback():
    datasets = pytest.importorskip("datasets")
    from viquae_b200.ingest import arrow_column_blocks
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 24)).astype(np.float32)
    ds = datasets.Dataset.from_dict({"emb": x.tolist(), "other": list(range(1000))})
    ds = ds.cast_column("emb", datasets.Sequence(datasets.Value("float32")))
    got = np.concatenate(list(arrow_column_blocks(ds, "emb", max_rows=300)))
    assert got.dtype == np.float32 and np.array_equal(got, x)
    # a selection (indices mapping) and ragged rows go through the formatted fallback
    sub = ds.select([5, 3, 900])
    assert np.array_equal(np.concatenate(list(arrow_column_blocks(sub, "emb"))), x[[5, 3, 900]])
    ds64 = datasets.Dataset.from_dict({"emb": x.astype(np.float64).tolist()})
    assert np.allclose(np.concatenate(list(arrow_column_blocks(ds64, "emb"))), x)


def test_faiss_flat_file_layout(tmp_path):
    """The FAISS "Flat" / "L2norm,Flat" file layout restated in viquae_b200/faiss_io.py (faiss is absent: the
    expected bytes are assembled by hand from the layout table in that module's header — parity unpinned)."""
    import struct
    from viquae_b200 import faiss_io
    rows = np.arange(12, dtype=np.float32).reshape(3, 4) / 7
    header = struct.pack("<i", 4) + struct.pack("<q", 3) + struct.pack("<q", 1 << 20) * 2 + b"\x01" + struct.pack("<i", 0)
    flat = b"IxFI" + header + struct.pack("<Q", 12) + rows.tobytes()
    p = tmp_path / "flat.faiss"
    faiss_io.write_flat_index(p, rows)
    assert p.read_bytes() == flat
    got = faiss_io.read_flat_index(p)
    assert got.l2norm is False and got.metric_type == 0 and got.string_factory == "Flat"
    np.testing.assert_array_equal(got.rows, rows)

    pre = b"IxPT" + header + struct.pack("<i", 1) + b"VNrm" + struct.pack("<f", 2.0) + struct.pack("<ii", 4, 4) + b"\x01" + flat
    p2 = tmp_path / "l2norm_flat.faiss"
    faiss_io.write_flat_index(p2, rows, l2norm=True)
    assert p2.read_bytes() == pre
    got = faiss_io.read_flat_index(p2)
    assert got.l2norm is True and got.string_factory == "L2norm,Flat"
    np.testing.assert_array_equal(got.rows, rows)

    # empty index, L2 metric fourcc, and the failure modes
    faiss_io.write_flat_index(tmp_path / "empty", np.zeros((0, 8), np.float32))
    assert faiss_io.read_flat_index(tmp_path / "empty").rows.shape == (0, 8)
    faiss_io.write_flat_index(tmp_path / "l2", rows, metric_type=1)
    assert (tmp_path / "l2").read_bytes()[:4] == b"IxF2" and faiss_io.read_flat_index(tmp_path / "l2").metric_type == 1
    (tmp_path / "ivf").write_bytes(b"IwFl" + header)
    with pytest.raises(NotImplementedError):
        faiss_io.read_flat_index(tmp_path / "ivf")
    (tmp_path / "cut").write_bytes(flat[:-5])
    with pytest.raises(ValueError, match="truncated"):
        faiss_io.read_flat_index(tmp_path / "cut")
    (tmp_path / "bad").write_bytes(b"IxFI" + header + struct.pack("<Q", 11) + rows.tobytes())
    with pytest.raises(ValueError):
        faiss_io.read_flat_index(tmp_path / "bad")
    (tmp_path / "pca").write_bytes(b"IxPT" + header + struct.pack("<i", 1) + b"PcAm")
    with pytest.raises(NotImplementedError):
        faiss_io.read_flat_index(tmp_path / "pca")


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny workload: one JSON line with
    the contract's keys, no GPU needed."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--rows", "20000", "--nq", "8",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] and "model" not in line["config"]


def test_resolve_devices_follows_datasets_device_semantics():
    """`device` of an index config (datasets/search.py:316-347: None, int >= 0, negative = all GPUs, list = those
    GPUs) -> (this process's GPU, sharded?) for the one-process-per-GPU engine."""
    from viquae_b200.sharded import resolve_devices, shard_range
    assert resolve_devices(None, 1, 0, 8) == (0, False)
    assert resolve_devices(3, 1, 0, 8) == (3, False)
    assert resolve_devices(None, 8, 5, 8) == (5, False)          # a replica per rank, like FAISS replicas
    assert resolve_devices(-1, 8, 5, 8) == (5, True)
    assert resolve_devices(-1, 1, 0, 1) == (0, False)            # a 1-GPU box: "all GPUs" is that GPU
    assert resolve_devices([4, 6], 2, 1, 8) == (6, True)
    assert resolve_devices([2], 1, 0, 8) == (2, False)
    with pytest.raises(RuntimeError, match="nproc-per-node 8"):
        resolve_devices(-1, 1, 0, 8)                             # all GPUs from one process: tell the user how to launch
    with pytest.raises(RuntimeError, match="nproc-per-node 2"):
        resolve_devices([0, 1], 4, 0, 8)
    with pytest.raises(ValueError):
        resolve_devices([0, 0], 2, 0, 8)
    # the row ranges of the ranks tile the corpus
    cuts = [shard_range(1_000_003, r, 8) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 1_000_003 and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


def test_committed_bench_lines_keep_the_contract():
    """Every bench line kept under profiles/ as round-2 evidence still carries the keys the driver's contract names:
    metric/value/unit, the roofline and e2e objects, a launch count, and (where the run checked it) a green parity check;
    lines of one corpus measured at different GPU counts share one result digest."""
    import glob
    import json
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
    files = sorted(glob.glob(os.path.join(root, "r2_bench_*.json")))
    assert files, "no round-2 bench lines committed"
    digests = {}
    for f in files:
        line = json.loads(open(f).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "dtype", "data", "config", "e2e"):
            assert key in line, (os.path.basename(f), key)
        assert line["value"] > 0 and line["e2e"]["value"] > 0
        if line.get("impl") == "reference":
            assert line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0
            continue
        assert line["gpu_launches"] > 0, os.path.basename(f)
        roof = line["roofline"]
        assert roof["bound"] in ("hbm", "tensor", "fp32-ffma") and 0 < roof["frac"] < 1.2 and roof["peak"] > 0, os.path.basename(f)
        if line.get("parity_check"):
            assert line["parity_check"]["ok"] is True, os.path.basename(f)
        if "final_tree" in f and "config5" not in f:
            digests[line["n_gpus"]] = line["result_sha"]
    assert len(digests) >= 2 and len(set(digests.values())) == 1, digests

