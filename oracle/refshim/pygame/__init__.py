"""Stub of pygame: the reference only reads K_* constants at import time."""


def init():
    return None


def __getattr__(name):
    if name.startswith("K_"):
        return sum(ord(c) for c in name)
    raise AttributeError(name)
