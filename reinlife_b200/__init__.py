"""reinlife_b200 -- B200-native (sm_100a) implementation of ReinLife's data-parallel hot path.

Mirrors the reference's import surface (ReinLife/__init__.py:1-5):
    from reinlife_b200 import trainer, tester, Environment, Models, Saver
The CUDA library is the product: anything that touches the device raises if libreinlife_b200.so is missing.
"""
from . import Models                      # noqa: F401
from .Helpers.saver import Saver          # noqa: F401
from .Helpers.trainer import trainer      # noqa: F401
from .Helpers.tester import tester        # noqa: F401
from .World.environment import Environment  # noqa: F401
