"""CPU: table compiler, layout sampler, C-ABI symbol table (no GPU calls)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

from cooking_zoo_b200 import _native, levels
from cooking_zoo_b200.layout import sample_layout
from cooking_zoo_b200.tables import compile_tables
from tests.replay import golden_files, load_golden, ROOT


def test_library_exports_every_declared_symbol():
    """every function include/cz_b200.h declares is exported by libcz_b200.so and bound in _native"""
    header = open(os.path.join(ROOT, "include", "cz_b200.h")).read()
    declared = set(re.findall(r"\b(cz_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _native.load_library()
    assert lib.cz_abi_version() == _native.ABI_VERSION
    assert lib.cz_layout_draw(1, 2, 3) == lib.cz_layout_draw(1, 2, 3) != lib.cz_layout_draw(1, 2, 4)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cooking_zoo_b200 import BatchedCookingEnv
    with pytest.raises(_native.NativeError):
        BatchedCookingEnv(4, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], action_scheme="scheme3")


def test_sampler_reproduces_golden_layouts():
    """sample_layout(random.Random(seed)) == the world the reference built after random.seed(seed)
    (constructor + reset both consume the stream: the recorded layout is the second draw)."""
    g = load_golden(os.path.join(ROOT, "tests", "golden", "cfg2_uniform.npz"))
    cfg = g["config"]
    for seed, want in enumerate(g["layouts"]):
        rng = random.Random(seed)
        lo, meta = levels.load_level_object(cfg["level"]), levels.load_meta(cfg["meta_file"])
        sample_layout(lo, meta, cfg["num_agents"], rng)
        assert sample_layout(lo, meta, cfg["num_agents"], rng) == want


def test_compiled_tables_shapes_and_plan():
    t = compile_tables("coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"],
                       end_condition_all_dishes=True)
    assert t.obs_len == 278 and t.num_canon_slots == 28 and t.num_dyn_slots == 12
    assert t.num_variants == 2 and t.rows == 12 + 2 + 6
    assert t.obs_segs[:t.num_obs_segs].tolist() == [[0, 108, 0], [262, 16, 108]]
    assert t.obs_ranges[:t.num_obs_ranges].tolist() == [[108, 154]]
    assert t.num_comp_slots == 14            # 12 live dynamic slots + 2 agents
    assert t.reward_time == -5 / 400
    # the static table holds (sx - ax) / W exactly as Python divides
    cell = 2 * 8 + 1                          # observer at (1, 2)
    sc = int(t.static_cells[0, 0])
    assert t.obs_table[0, cell, 0] == ((sc & 7) - 1) / 7 and t.obs_table[0, cell, 2] == 1.0


def test_compile_rejects_what_the_reference_rejects():
    with pytest.raises(AssertionError):       # cooking_env.py:93-94
        compile_tables("coop_test", "example", 3, 400, ["TomatoSalad"] * 3)
    with pytest.raises(ValueError):           # compute_infos would raise IndexError (cooking_env.py:329)
        compile_tables("coop_test", "example", 2, 400, ["TomatoSalad"])
    with pytest.raises(FileNotFoundError):
        compile_tables("no_such_level", "example", 1, 400, ["TomatoSalad"])


@pytest.mark.parametrize("path", golden_files()[:4], ids=lambda p: p.split("/")[-1][:-4])
def test_tables_compile_for_golden_configs(path):
    from tests.replay import package_recipes
    g = load_golden(path)
    cfg = g["config"]
    with package_recipes(cfg):
        t = compile_tables(cfg["level"], cfg["meta_file"], cfg["num_agents"], cfg["max_steps"], cfg["recipes"],
                           cfg["reward_scheme"], cfg["end_all"], layouts=g["layouts"])
    assert t.num_layouts == len(g["layouts"])
    # pooled initial records decode to the recorded initial object positions
    objs = g["objs"][0, 0]
    dev = t.pool[0, :t.num_dyn_slots]
    for d, c in enumerate(t.canon_of_dev):
        if objs[c, 0]:
            assert (int(dev[d]) & 7, (int(dev[d]) >> 3) & 7) == (objs[c, 1], objs[c, 2])


def test_planner_keeps_the_shipped_levels_in_the_specialised_class():
    """what cz_tables_create needs for the specialised kernels + pipelined mode: few variants, one computed
    range, at most 64 (observer, slot) pairs, even rows, table run <= 128 doubles"""
    rec = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
    cases = [("coop_test", "example", 2), ("switch_test", "example", 2), ("coexistence_test", "example", 2)]
    cases += [(os.path.join(ROOT, "tests/golden/levels/open4.json"), os.path.join(ROOT, "tests/golden/levels/meta4.json"), a)
              for a in (1, 2, 3, 4)]
    for level, meta, agents in cases:
        t = compile_tables(level, meta, agents, 100, rec[:agents], layout_pool_size=300, layout_seed=1)
        assert t.num_variants <= 16, (level, t.num_variants)
        assert t.num_obs_ranges == 1 and t.obs_len % 2 == 0 and t.obs_table_len <= 128, level
        assert 0 < agents * t.num_comp_slots <= 64, (level, agents)
        # rows are partitioned into table segments and the computed range
        covered = np.zeros(t.obs_len, int)
        for k in range(t.num_obs_segs):
            covered[t.obs_segs[k, 0]:t.obs_segs[k, 0] + t.obs_segs[k, 1]] += 1
        covered[t.obs_ranges[0, 0]:t.obs_ranges[0, 0] + t.obs_ranges[0, 1]] += 1
        assert (covered == 1).all(), level


def test_scan_order_is_shared_only_when_every_layout_agrees_with_the_level_file():
    """OPTIONAL objects: layouts whose type insertion order is a subsequence of the level file's order share one scan order
    (variants do not multiply); a level whose optional first entry can move a type behind another keeps per-layout orders"""
    t = compile_tables("coexistence_test", "example", 2, 100, ["TomatoLettuceSalad", "CarrotBanana"], layout_pool_size=200)
    assert len({s.tobytes() for s in t.scan_order}) == 1
    lvl = dict(levels.load_level_object("coexistence_test"))
    # Tomato may appear before or after Lettuce: first entry optional, second entry later in the file
    dyn = [dict(e) for e in lvl["DYNAMIC_OBJECTS"]]
    tomato = next(e for e in dyn if "Tomato" in e)
    early = {"Tomato": dict(tomato["Tomato"], OPTIONAL=0.5)}
    late = {"Tomato": dict(tomato["Tomato"], OPTIONAL=1.0, X_POSITION=[6], Y_POSITION=[5])}
    dyn = [early] + [e for e in dyn if "Tomato" not in e] + [late]
    lvl["DYNAMIC_OBJECTS"] = dyn
    import json
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(lvl, f)
    try:
        t2 = compile_tables(f.name, "example", 2, 100, ["TomatoLettuceSalad", "CarrotBanana"], layout_pool_size=200)
    finally:
        os.unlink(f.name)
    assert len({s.tobytes() for s in t2.scan_order}) > 1


def _chi2_ok(counts, probs, n):
    """Pearson chi-square of observed counts against exact probabilities: within 5 sigma of its mean"""
    import numpy as np
    exp = np.asarray(probs, np.float64) * n
    chi2 = float(((np.asarray(counts) - exp) ** 2 / exp).sum())
    dof = len(probs) - 1
    return abs(chi2 - dof) < 5 * (2 * dof) ** 0.5, chi2, dof


def test_exact_layout_distribution_matches_the_parser():
    """layout.enumerate_layouts: the support and the probabilities of the reference's level parser, against the
    frequencies of sample_layout (which reproduces the reference's worlds seed by seed, see above)"""
    from fractions import Fraction
    from cooking_zoo_b200.layout import enumerate_layouts
    for level, meta, A, n_support in (("coop_test", "example", 2, 400), ("switch_test", "example", 2, 648),
                                      (os.path.join(ROOT, "tests/golden/levels/weighted_layouts.json"), "example", 1, None)):
        lo, mt = levels.load_level_object(level), levels.load_meta(meta)
        support = enumerate_layouts(lo, mt, A)
        assert sum(p for _, p in support) == 1
        if n_support:
            assert len(support) == n_support and all(p == Fraction(1, n_support) for _, p in support)   # 400
        else:
            assert len({p for _, p in support}) > 3          # OPTIONAL objects, duplicate list entries, competing objects
        index = {repr(lay): k for k, (lay, _) in enumerate(support)}
        n = 60 * len(support)
        counts = [0] * len(support)
        rng = random.Random(99)
        for _ in range(n):
            counts[index[repr(sample_layout(lo, mt, A, rng))]] += 1       # KeyError = a layout outside the support
        ok, chi2, dof = _chi2_ok(counts, [float(p) for _, p in support], n)
        assert ok, (level, chi2, dof)


def test_auto_pool_is_the_exact_support_and_large_levels_fall_back():
    t = compile_tables("coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"])
    assert t.layout_exact and t.num_layouts == 400 and t.layout_cum is not None
    assert int(t.layout_cum[-1]) == 2 ** 64 - 1 and (np.diff(t.layout_cum.astype(object)) > 0).all()
    t = compile_tables("coexistence_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"])
    assert not t.layout_exact and t.num_layouts == 256 and t.layout_cum is None      # > 4096 layouts: i.i.d. pool
    t = compile_tables("coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], layout_pool_size=32)
    assert not t.layout_exact and t.num_layouts == 32 and t.layout_cum is None
