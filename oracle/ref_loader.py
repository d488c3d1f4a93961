"""Locate and import the UNMODIFIED reference (test / benchmark infrastructure).

The reference lives read-only at /root/reference in the build container; `python -m oracle.make_ref` copies its
package verbatim (SHA-256 manifest) into the git-ignored oracle/_ref/, which travels to the GPU box with the
snapshot, so `bench.py --impl reference` and the live lockstep tests can run there too.  It
needs gymnasium / pettingzoo / pygame at import time (cooking_zoo/__init__.py:1,
environment/cooking_env.py:11-17, environment/game/graphic_pipeline.py:6); none of them is
installed and none performs hot-path arithmetic, so `oracle/refshim/` supplies import
stubs.  Only tests/golden/make_golden.py and the live cross-check
tests call this; the product package never does.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "refshim")


def _default_root():
    if os.path.isdir("/root/reference/cooking_zoo"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")          # the verbatim copy made by oracle/make_ref.py


REFERENCE_ROOT = os.environ.get("CZ_REFERENCE_ROOT") or _default_root()


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
