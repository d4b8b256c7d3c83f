"""unimp_b200 — B200-native (sm_100a) implementation of UniMP's OpenFlamingo hot path.

Public surface mirrors what the reference imports from `open_flamingo`
(`UniMP/mmrec.py:20-22`): `create_model_and_transforms`, `Flamingo`.
Importing this package does not need a GPU; running any op does (no CPU fallback).
"""
from .config import FlamingoConfig, openflamingo_4b_config, tiny_config, WORKLOADS  # noqa: F401

__all__ = ["create_model_and_transforms", "Flamingo", "build_flamingo"]


def __getattr__(name):
    if name in ("create_model_and_transforms", "build_flamingo", "SyntheticTokenizer"):
        from . import factory
        return getattr(factory, name)
    if name == "Flamingo":
        from .flamingo import Flamingo
        return Flamingo
    raise AttributeError(name)
