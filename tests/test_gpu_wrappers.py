"""GPU: the reference's wrapper surface (parallel_env / gym single / gym multi-agent) over the CUDA backend,
replayed against golden traces recorded from the reference."""
import os

import numpy as np
import pytest

from tests.replay import load_golden, bits, ROOT

pytestmark = pytest.mark.gpu


def _kw(cfg, layouts):
    return dict(level=cfg["level"], meta_file=cfg["meta_file"], max_steps=cfg["max_steps"], recipes=cfg["recipes"],
                obs_spaces=["feature_vector"] * cfg["num_agents"], end_condition_all_dishes=cfg["end_all"],
                action_scheme="scheme3", reward_scheme=cfg["reward_scheme"], layouts=layouts)


def test_parallel_env_dict_api_matches_reference_trace():
    from cooking_zoo_b200.wrappers import parallel_env
    g = load_golden(os.path.join(ROOT, "tests", "golden", "heuristic_any.npz"))
    cfg = g["config"]
    env = parallel_env(num_agents=cfg["num_agents"], **_kw(cfg, g["layouts"]))
    assert env.possible_agents == ["player_0", "player_1"]
    assert env.observation_space("player_0").shape == (278,) and env.action_space("player_1").n == 5
    for n in range(2):
        obs, infos = env.reset(options={"layout_id": n})
        assert set(obs) == {"player_0", "player_1"} and infos["player_0"] == {}
        for i in range(2):
            assert np.array_equal(bits(g["obs"][n, 0, i]), bits(obs[f"player_{i}"]))
        for t in range(int(g["length"][n])):
            act = {f"player_{i}": int(g["actions"][n, t, i]) for i in range(2)}
            obs, rew, term, trunc, infos = env.step(act)
            for i in range(2):
                a = f"player_{i}"
                assert np.array_equal(bits(g["obs"][n, t + 1, i]), bits(obs[a])), (n, t)
                assert bits(rew[a]) == bits(g["reward"][n, t, i]) and isinstance(rew[a], np.float64)
                assert term[a] == bool(g["term"][n, t, i]) and trunc[a] == bool(g["trunc"][n, t, i])
                assert infos[a]["t"] == t + 1 and infos[a]["task"] == cfg["recipes"][i]
                assert infos[a]["action"] == act[a] and infos[a]["goal_vector"][i] == 1.0
        assert any(term.values()) or any(trunc.values())
        with pytest.raises(RuntimeError):
            env.step({"player_0": 0, "player_1": 0})
    env.close()


def test_gym_single_and_multi_agent_shapes():
    from cooking_zoo_b200.wrappers import GymCookingEnvironment, GymCookingEnvironmentMA
    g1 = load_golden(os.path.join(ROOT, "tests", "golden", "heuristic_cfg1.npz"))
    cfg = g1["config"]
    env = GymCookingEnvironment(**_kw(cfg, g1["layouts"]))
    obs, info = env.reset(options={"layout_id": 0})
    assert obs.shape == (278,) and obs.dtype == np.float64
    assert np.array_equal(bits(g1["obs"][0, 0, 0]), bits(obs))
    total = 0.0
    for t in range(int(g1["length"][0])):
        obs, r, term, trunc, info = env.step(int(g1["actions"][0, t, 0]))
        assert np.array_equal(bits(g1["obs"][0, t + 1, 0]), bits(obs))
        assert bits(r) == bits(g1["reward"][0, t, 0])
        total += r
    assert term and not trunc and info["recipe_done"]          # the scripted cook finishes the salad
    env.close()
    g2 = load_golden(os.path.join(ROOT, "tests", "golden", "cfg2_trunc25.npz"))
    cfg = g2["config"]
    ma = GymCookingEnvironmentMA(num_agents=2, **_kw(cfg, g2["layouts"]))
    obs, infos = ma.reset(options={"layout_id": 1})
    assert isinstance(obs, list) and len(obs) == 2 and len(infos) == 2
    for t in range(int(g2["length"][1])):
        obs, rew, term, trunc, infos = ma.step([int(a) for a in g2["actions"][1, t]])
        assert np.array_equal(bits(g2["obs"][1, t + 1]), bits(np.stack(obs)))
    assert trunc == [True, True] and term == [False, False]
    assert infos[0]["termination_info"] == "Terminating because 25 timesteps passed"
    ma.close()
