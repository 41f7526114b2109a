"""Loads the reference's own meerqat.ir.search / meerqat.ir.fuse from /root/reference, unmodified,
behind sys.modules stubs for the third-party packages that are not installable offline (docopt,
ranx, jsonargparse, elasticsearch, pyserini) and a faiss API shim (tests/shims/faiss).
Only usable where /root/reference exists (this container), i.e. for generating golden fixtures and
for the `reference`-marked tests; nothing that runs on the GPU box may import this."""
import importlib
import os
import sys
import types

REFERENCE = os.environ.get("VIQUAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REFERENCE, "meerqat"))


class _Run:
    """Minimal ranx.Run: name + dict q -> {doc -> score} sorted by descending score."""

    def __init__(self, run=None, name=None):
        self.name = name
        self.run = {}
        for q, res in (run or {}).items():
            self.run[q] = dict(sorted(res.items(), key=lambda kv: -kv[1]))

    def keys(self):
        return self.run.keys()

    def __getitem__(self, q):
        return self.run[q]

    def save(self, path):
        import json
        with open(path, "wt") as f:
            json.dump(self.run, f)


class _Qrels:
    def __init__(self, qrels=None, name=None):
        self.qrels = dict(qrels or {})

    def save(self, path):
        import json
        with open(path, "wt") as f:
            json.dump(self.qrels, f)


def install_stubs():
    shim_dir = os.path.join(HERE, "shims")
    if shim_dir not in sys.path:
        sys.path.insert(0, shim_dir)          # `import faiss` -> the shim; must precede `import datasets`
    import faiss  # noqa: F401
    import datasets
    if not hasattr(datasets, "set_caching_enabled"):  # removed in datasets>=3, imported at search.py:33
        datasets.set_caching_enabled = lambda flag: (datasets.enable_caching() if flag else datasets.disable_caching())
    import datasets.search as ds_search
    if not getattr(ds_search, "_has_faiss", False):
        ds_search._has_faiss = True

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
        return sys.modules[name]

    stub("docopt", docopt=lambda *a, **k: {})
    stub("ranx", Run=_Run, Qrels=_Qrels, fuse=None, optimize_fusion=None, evaluate=None, compare=None)
    stub("jsonargparse", CLI=lambda *a, **k: None)

    class _ES:
        def __init__(self, *a, **k):
            pass
    stub("elasticsearch", Elasticsearch=_ES)
    # heavy siblings of the reference package that this path only references by name
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    pkg = stub("meerqat")
    pkg.__path__ = [os.path.join(REFERENCE, "meerqat")]
    ir = stub("meerqat.ir")
    ir.__path__ = [os.path.join(REFERENCE, "meerqat", "ir")]
    data = stub("meerqat.data")
    data.__path__ = [os.path.join(REFERENCE, "meerqat", "data")]
    stub("meerqat.ir.metrics", find_relevant=lambda retrieved, *a, **k: ([], []), fuse_qrels=lambda paths: None)
    import enum

    class QuestionType(enum.Enum):
        String = 0
        Numerical = 1
        Time = 2
    stub("meerqat.data.infoseek", QuestionType=QuestionType)
    stub("meerqat.data.utils", json_integer_keys=lambda d: {int(k): v for k, v in d.items()}, to_latex=lambda m: str(m))


def load_reference_search():
    install_stubs()
    return importlib.import_module("meerqat.ir.search")


def load_reference_fuse():
    install_stubs()
    return importlib.import_module("meerqat.ir.fuse")
