"""CPU: the compiled oracle (oracle/cz_oracle.c) replayed against the golden traces recorded from the reference."""
import numpy as np
import pytest

from oracle.cz_oracle import SpawnStream
from oracle.cz_oracle_c import COracleEnv
from tests.replay import golden_files, load_golden, assert_state_equal, assert_obs_equal, bits, STATE_KEYS


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_c_oracle_replays_golden(path):
    g = load_golden(path)
    cfg = g["config"]
    A = cfg["num_agents"]
    for n, layout in enumerate(g["layouts"]):
        sp = cfg.get("spawn")
        kw = {} if not sp else dict(agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                                    grace_period=sp["grace"], spawn_stream=SpawnStream(sp["seed"], n, 1))
        env = COracleEnv(layout, cfg["recipes"], cfg["max_steps"], reward_scheme=cfg["reward_scheme"],
                         end_condition_all_dishes=cfg["end_all"], action_scheme=cfg.get("action_scheme", "scheme3"), **kw)
        ctx = f"{path} trace {n} reset"
        assert_state_equal({k: g[k][n, 0] for k in STATE_KEYS}, env.export_state(), ctx)
        assert_obs_equal(g["obs"][n, 0], np.stack([env.observe(i) for i in range(A)]), ctx)
        for t in range(int(g["length"][n])):
            ctx = f"{path} trace {n} step {t}"
            for i in range(A):
                if g["teleport"][n, t, i, 0] >= 0:
                    env.teleport(i, *map(int, g["teleport"][n, t, i]))
            rew, term, trunc, rel = env.step(g["actions"][n, t])
            assert np.array_equal(bits(g["reward"][n, t]), bits(rew)), ctx
            assert np.array_equal(g["term"][n, t], term) and np.array_equal(g["trunc"][n, t], trunc), ctx
            assert np.array_equal(g["rel"][n, t], rel), ctx
            assert_state_equal({k: g[k][n, t + 1] for k in STATE_KEYS}, env.export_state(), ctx)
            assert_obs_equal(g["obs"][n, t + 1], np.stack([env.observe(i) for i in range(A)]), ctx)
        assert env.error == 0
