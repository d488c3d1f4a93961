"""Stub of gymnasium: only what `import cooking_zoo` touches (spaces, Env, register)."""
from . import spaces, envs, utils  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass
