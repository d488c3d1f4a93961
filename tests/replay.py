"""Shared lockstep-replay helpers: golden trace / live env  vs  an env under test."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_KEYS = ("agents", "objs", "statics", "marks")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tree(t):
    return (t[0], t[1], [_tree(k) for k in t[2]])


def load_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["config"] = json.loads(str(g["config"]))
    g["layouts"] = json.loads(str(g["layouts"]))
    # recipes registered through the reference's register_recipe hook: make the oracle know them (additive)
    from oracle.cz_oracle import register_recipe
    for name, tree in g["config"].get("custom_recipes", {}).items():
        register_recipe(name, _tree(tree))
    for key in ("level", "meta_file"):        # repo-relative .json paths (custom levels)
        if g["config"][key].endswith(".json"):
            g["config"][key] = os.path.join(ROOT, g["config"][key])
    return g


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_state_equal(want, got, ctx):
    for k in STATE_KEYS:
        w, g = np.asarray(want[k]), np.asarray(got[k])
        if not np.array_equal(w, g):
            bad = np.argwhere(w != g)
            raise AssertionError(f"{ctx}: state[{k}] differs at {bad[:8].tolist()}\nwant={w[tuple(bad[0][:-1])] if w.ndim > 1 else w}\n"
                                 f"got ={g[tuple(bad[0][:-1])] if g.ndim > 1 else g}")


def assert_obs_equal(want, got, ctx):
    wb, gb = bits(want), bits(got)
    if not np.array_equal(wb, gb):
        bad = np.argwhere(wb != gb)
        i = tuple(bad[0])
        raise AssertionError(f"{ctx}: obs differs at {bad[:8].tolist()} want {np.asarray(want)[i]!r} got {np.asarray(got)[i]!r}")


import contextlib


@contextlib.contextmanager
def package_recipes(cfg):
    """Register cfg["custom_recipes"] with the product's register_recipe (which, like the reference's, makes the store
    replace the book: cooking_env.py:100-105) for the duration of a test."""
    custom = cfg.get("custom_recipes")
    if not custom:
        yield
        return
    from cooking_zoo_b200 import recipes as R
    conds = {"chopped": [("chop_state", "Chopped")], "mashed": [("blend_state", "Mashed")], None: None}

    def build(t):
        return R.RecipeNode(name=t[0], conditions=conds[t[1]], contains=[build(k) for k in t[2]])
    saved = dict(R.RECIPE_STORE)
    R.RECIPE_STORE.clear()
    for name, tree in custom.items():
        R.register_recipe(R.Recipe(build(tree)), name)
    try:
        yield
    finally:
        R.RECIPE_STORE.clear()
        R.RECIPE_STORE.update(saved)
