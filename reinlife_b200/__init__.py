"""reinlife_b200 -- B200-native (sm_100a) implementation of ReinLife's data-parallel hot path.

Mirrors the reference's import surface (ReinLife/__init__.py:1-5):
    from reinlife_b200 import trainer, tester, Environment, Models
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing when first used)
