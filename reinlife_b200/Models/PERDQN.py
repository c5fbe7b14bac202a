"""PERDQN (ReinLife/Models/PERDQN.py) -- the reference's fifth exported brain.  SURVEY.md 8(f) rank 3: not on the
north-star path; not implemented yet.  Constructing it fails loudly instead of silently running somewhere else."""
from .utils import BasicBrain


class PERDQNAgent(BasicBrain):
    def __init__(self, *args, **kwargs):
        super().__init__(153, 8, "PERDQN")
        raise NotImplementedError("PERDQN (SumTree PER, 153-64-64-8) is not implemented in reinlife_b200 yet")
