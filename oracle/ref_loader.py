"""Locate and import the UNMODIFIED reference (test infrastructure, build container only).

The reference lives read-only at /root/reference and does not exist on the GPU box.  It
needs gymnasium / pettingzoo / pygame at import time (cooking_zoo/__init__.py:1,
environment/cooking_env.py:11-17, environment/game/graphic_pipeline.py:6); none of them is
installed and none performs hot-path arithmetic, so `oracle/refshim/` supplies import
stubs.  Only tests/golden/make_golden.py and the live cross-check
tests call this; the product package never does.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("CZ_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cooking_zoo"))


def load_reference():
    """Returns the reference's `cooking_env` module (CookingEnvironment lives there)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    from cooking_zoo.environment import cooking_env  # noqa: E402
    return cooking_env
