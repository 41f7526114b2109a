"""viquae_b200 — B200-native exact inner-product top-k search + score fusion behind the retrieval API of
PaulLerner/ViQuAE (``meerqat.ir.search`` / ``meerqat.ir.fuse``).

Host code is Python; every vector operation is a call into ``libviquae_b200.so`` (hand-written
sm_100a CUDA behind the C ABI in ``include/viquae_b200.h``).  There is no CPU or PyTorch fallback:
importing the engine modules without the built library raises.
"""
__version__ = "0.1.0"

__all__ = ["B200FlatIndex", "ShardedFlatIndex", "ShardedB200Index", "KnowledgeBase", "Searcher", "IndexKind", "Index", "Fusion", "Run", "Qrels",
           "inbatch_step", "mine_hard_negatives"]


def __getattr__(name):  # lazy: `import viquae_b200` stays cheap and works where datasets/torch are heavy
    if name == "B200FlatIndex":
        from .index import B200FlatIndex
        return B200FlatIndex
    if name == "ShardedFlatIndex":
        from .sharded import ShardedFlatIndex
        return ShardedFlatIndex
    if name in ("KnowledgeBase", "Searcher", "IndexKind", "Index", "dataset_search", "L2norm"):
        from . import search
        return getattr(search, name)
    if name in ("Fusion", "default_minimum", "gzmuv_norm"):
        from . import fuse
        return getattr(fuse, name)
    if name in ("inbatch_step", "mine_hard_negatives", "gather_global_batch"):
        from . import inbatch
        return getattr(inbatch, name)
    if name in ("ShardedB200Index", "PeerExchange"):
        from . import sharded
        return getattr(sharded, name)
    if name in ("Run", "Qrels", "evaluate"):
        from . import runs
        return getattr(runs, name)
    raise AttributeError(name)
