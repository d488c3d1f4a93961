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


def test_symbolic_observation_rebuilt_from_device_state_matches_the_oracle_world():
    """the "symbolic" view (world_objects by class name + Agent) of a device environment"""
    import numpy as np
    import torch
    from oracle.cz_oracle import OracleEnv
    from cooking_zoo_b200 import BatchedCookingEnv
    recipes = ["TomatoLettuceSalad", "CarrotBanana"]
    n = 6
    env = BatchedCookingEnv(n, "coop_test", "example", 2, 300, recipes, end_condition_all_dishes=True,
                            action_scheme="scheme3", layout_pool_size=8)
    lids = np.arange(n, dtype=np.int32)
    env.reset(layout_ids=lids)
    oracles = [OracleEnv(env.tables.layouts[l], recipes, 300, end_condition_all_dishes=True) for l in lids]

    for t in range(120):
        act, _ = env.heuristic_actions()           # the cook moves things around: plates get filled, food chopped
        env.step(act)
        a = act.cpu().numpy()
        for k, orc in enumerate(oracles):
            orc.step(a[k])
        if t % 10 != 9:
            continue
        for k, orc in enumerate(oracles):
            sym = env.symbolic_observation(k)
            for typ, lst in orc.by_type.items():
                got = sym.get(typ, [])
                assert [r.location for r in got] == [(o.x, o.y) for o in lst], (t, k, typ)
                for r, o in zip(got, lst):
                    assert [(c.name, c.location) for c in r.content] == [(c.type, (c.x, c.y)) for c in o.content], (t, k, typ)
                    if hasattr(r, "chop_state"):
                        assert (r.chop_state == "CHOPPED") == o.chopped and r.free == o.free
                    if hasattr(r, "blend_state"):
                        assert (r.blend_state == "MASHED") == (o.blend == 2)
            for r, ag in zip(sym["Agent"], orc.agents):
                assert r.location == (ag.x, ag.y) and r.orientation == ag.orientation
                assert (r.holding is None) == (ag.holding is None)
                if ag.holding is not None:
                    assert (r.holding.name, r.holding.location) == (ag.holding.type, (ag.holding.x, ag.holding.y))


def test_aec_surface_replays_the_reference_trace_call_by_call():
    """cooking_env.env(...): agent_selection / last() / per-agent step against tests/golden/aec/aec_cfg2.npz, recorded
    from the unmodified reference (make_golden.py aec), including the cumulative-reward quirk of cooking_env.py:228"""
    import json
    from cooking_zoo_b200.wrappers import env as make_env, ENV_IDS
    z = np.load(os.path.join(ROOT, "tests", "golden", "aec", "aec_cfg2.npz"))
    cfg, layouts = json.loads(str(z["config"])), json.loads(str(z["layouts"]))
    assert ENV_IDS["cookingZooEnv-v0"].endswith("AECCookingEnv")
    for k in range(len(layouts)):
        aec = make_env(level=cfg["level"], meta_file=cfg["meta_file"], num_agents=2, max_steps=cfg["max_steps"],
                       recipes=cfg["recipes"], obs_spaces=["feature_vector"] * 2, end_condition_all_dishes=cfg["end_all"],
                       action_scheme=cfg["action_scheme"], layouts=[layouts[k]])
        aec.reset(options={"layout_id": 0})
        n = int(z["n_calls"][k])
        for c in range(n):
            assert aec.agent_selection == f"player_{int(z['agent'][k, c])}", (k, c)
            obs, cum, term, trunc, info = aec.last()
            assert np.array_equal(bits(z["obs"][k, c]), bits(obs)), (k, c)
            assert bits(float(cum)) == bits(z["cum"][k, c]), (k, c, cum)
            assert (int(term), int(trunc)) == (int(z["term"][k, c]), int(z["trunc"][k, c])), (k, c)
            assert info.get("t", -1) == z["info_t"][k, c] and int(info.get("recipe_done", -1)) == z["info_done"][k, c]
            if z["action"][k, c] < 0:
                with pytest.raises(ValueError):
                    aec.step(0)
                aec.step(None)                      # C-8: a no-op, the round cannot make progress
                assert c == n - 1
            else:
                aec.step(int(z["action"][k, c]))
        aec.close()


def test_parallel_env_with_despawn_and_respawn_follows_the_reference_agents_list():
    """an agent that despawns leaves `agents` (truncated on that step) and the episode goes on for the others
    (cooking_env.py:262-266); replayed against spawn_cfg2.npz (rel = the agents present in the reference's dicts)"""
    from cooking_zoo_b200.wrappers import parallel_env
    g = load_golden(os.path.join(ROOT, "tests", "golden", "spawn_cfg2.npz"))
    cfg, sp = g["config"], g["config"]["spawn"]
    seen_partial = False
    for n in range(4):
        env = parallel_env(num_agents=2, agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                           grace_period=sp["grace"], seed=sp["seed"], env_offset=n, **_kw(cfg, [g["layouts"][n]]))
        env.reset(options={"layout_id": 0})
        for t in range(int(g["length"][n])):
            obs, rew, term, trunc, infos = env.step({f"player_{i}": int(g["actions"][n, t, i]) for i in range(2)})
            rel = [f"player_{i}" for i in range(2) if g["rel"][n, t, i]]
            assert sorted(obs) == rel and sorted(rew) == rel and env.agents == rel, (n, t)
            seen_partial |= len(rel) == 1
            for i in range(2):
                a = f"player_{i}"
                if a in rel:
                    assert np.array_equal(bits(g["obs"][n, t + 1, i]), bits(obs[a])) and bits(rew[a]) == bits(g["reward"][n, t, i])
                    assert term[a] == bool(g["term"][n, t, i]) and trunc[a] == bool(g["trunc"][n, t, i])
                    assert infos[a]["termination_info"] == ""
        env.close()
    assert seen_partial
