from .trainer import trainer  # noqa: F401
from .tester import tester    # noqa: F401
