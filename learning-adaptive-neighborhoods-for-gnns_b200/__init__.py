"""dgg_b200 -- B200-native (sm_100a) kernels for the Differentiable Graph Generator hot path.

The directory is named ``learning-adaptive-neighborhoods-for-gnns_b200``; ``dgg_b200.py`` at the repo
root registers it under the importable name ``dgg_b200``.
"""
from . import build  # noqa: F401
from ._lib import DggbError, declared_symbols, lib  # noqa: F401
from .graph import CSRGraph  # noqa: F401
from . import functional  # noqa: F401
from . import sharding  # noqa: F401
from .graphed import GraphedStep  # noqa: F401
from . import checkpoint  # noqa: F401
from . import losses  # noqa: F401
from . import samplers  # noqa: F401

__all__ = ["CSRGraph", "functional", "sharding", "GraphedStep", "lib", "build", "DggbError", "declared_symbols"]
