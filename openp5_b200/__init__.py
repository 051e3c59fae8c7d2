"""openp5_b200 — B200-native (sm_100a) engine for the OpenP5 T5 train step and trie-constrained beam search.

Public API: `P5B200` (model object with the reference's call surface), `Trie` (device trie),
`B200Runner` (drop-in for the reference's DistributedRunner), `build()` (compile libp5b200.so).
"""
from ._lib import P5LibraryError, LIB_PATH  # noqa: F401


def build(verbose: bool = False, force: bool = False) -> str:
    from .build import build as _b
    return _b(verbose=verbose, force=force)


def __getattr__(name):
    if name in ("P5B200", "Trie", "BACKBONES"):
        from . import model
        return getattr(model, name)
    if name == "B200Runner":
        from .runner import B200Runner
        return B200Runner
    raise AttributeError(name)
