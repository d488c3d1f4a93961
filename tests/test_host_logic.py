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
    g = load_golden(path)
    cfg = g["config"]
    t = compile_tables(cfg["level"], cfg["meta_file"], cfg["num_agents"], cfg["max_steps"], cfg["recipes"],
                       cfg["reward_scheme"], cfg["end_all"], layouts=g["layouts"])
    assert t.num_layouts == len(g["layouts"])
    # pooled initial records decode to the recorded initial object positions
    objs = g["objs"][0, 0]
    dev = t.pool[0, :t.num_dyn_slots]
    for d, c in enumerate(t.canon_of_dev):
        if objs[c, 0]:
            assert (int(dev[d]) & 7, (int(dev[d]) >> 3) & 7) == (objs[c, 1], objs[c, 2])
