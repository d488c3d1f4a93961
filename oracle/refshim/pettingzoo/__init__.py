"""Stub of pettingzoo: AECEnv base + the few utils cooking_env.py imports."""
from . import utils  # noqa: F401


class AECEnv:
    def __init__(self, *a, **k):
        pass

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def unwrapped(self):
        return self

    def _was_dead_step(self, action):
        raise NotImplementedError("stub: dead-step bookkeeping is not part of the hot path")
