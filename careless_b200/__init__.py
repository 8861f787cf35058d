"""careless_b200 -- B200-native ELBO gradient + Adam step of careless's VariationalMergingModel.

The numerics live in ``libcareless_b200.so`` (hand-written sm_100a CUDA, C-ABI in
``include/careless_b200.h``); this package is the Python host side that mirrors the
reference's Prior / Likelihood / Scaler / surrogate classes (``careless_b200.models``).
There is no CPU fallback: stepping a model without the built library or without a B200
raises.
"""
__version__ = "0.1.0"

from ._lib import ClbError, LibraryNotBuilt  # noqa: F401
from .engine import Engine, EngineConfig  # noqa: F401
