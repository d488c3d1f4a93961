"""CPU: oracle/cz_oracle.py replayed against the golden traces recorded from the reference."""
import numpy as np
import pytest

from oracle.cz_oracle import OracleEnv, NUM_GOALS, RECIPES, make_recipe, SpawnStream, spawn_uniform
from tests.replay import (golden_files, load_golden, assert_state_equal, assert_obs_equal, bits)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_replays_golden(path):
    g = load_golden(path)
    cfg = g["config"]
    A = cfg["num_agents"]
    for n, layout in enumerate(g["layouts"]):
        sp = cfg.get("spawn")
        kw = {} if not sp else dict(agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                                    grace_period=sp["grace"], spawn_stream=SpawnStream(sp["seed"], n, 1))
        env = OracleEnv(layout, cfg["recipes"], cfg["max_steps"], reward_scheme=cfg["reward_scheme"],
                        end_condition_all_dishes=cfg["end_all"], action_scheme=cfg.get("action_scheme", "scheme3"), **kw)
        ctx = f"{path} trace {n} reset"
        assert_state_equal({k: g[k][n, 0] for k in ("agents", "objs", "statics", "marks")}, env.export_state(), ctx)
        assert_obs_equal(g["obs"][n, 0], np.stack([env.observe(i) for i in range(A)]), ctx)
        for t in range(int(g["length"][n])):
            ctx = f"{path} trace {n} step {t}"
            for i in range(A):
                if g["teleport"][n, t, i, 0] >= 0:
                    env.teleport(i, *map(int, g["teleport"][n, t, i]))
            rew, term, trunc, rel = env.step(g["actions"][n, t])
            assert np.array_equal(bits(g["reward"][n, t]), bits(rew)), ctx
            assert list(g["term"][n, t]) == [int(v) for v in term], ctx
            assert list(g["trunc"][n, t]) == [int(v) for v in trunc], ctx
            assert list(g["rel"][n, t]) == [int(v) for v in rel], ctx
            assert_state_equal({k: g[k][n, t + 1] for k in ("agents", "objs", "statics", "marks")},
                               env.export_state(), ctx)
            assert_obs_equal(g["obs"][n, t + 1], np.stack([env.observe(i) for i in range(A)]), ctx)
        assert env.error == 0


def test_book_spot_values():
    """Survey-time spot values of the recipe book."""
    assert NUM_GOALS == 26
    assert len(RECIPES) == 8
    tls = make_recipe("TomatoLettuceSalad")
    assert [n.id for n in tls] == [18, 11, 2, 0]
    cb = make_recipe("CarrotBanana")
    assert [n.id for n in cb] == [20, 13, 8, 6]


def test_spawn_stream_matches_the_library():
    """oracle.spawn_uniform restates cz_spawn_uniform (include/cz_b200.h) bit for bit"""
    from cooking_zoo_b200 import _native
    lib = _native.load_library()
    for args in [(0, 0, 0, 0, 0), (4242, 3, 1, 17, 2), (2**63 + 5, 10**6, 40, 399, 1001)]:
        assert lib.cz_spawn_uniform(*args) == spawn_uniform(*args)
        assert 0.0 <= spawn_uniform(*args) < 1.0
