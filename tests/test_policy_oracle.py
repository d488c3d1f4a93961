"""CPU: oracle/cz_policy.py (the scripted cook restated) against the raw CookingAgent decisions
recorded from the unmodified reference (tests/golden/policy_*.npz), plus the host-built
reachability / first-step tables the device policy reads."""
import glob
import os

import numpy as np
import pytest

from oracle.cz_oracle import OracleEnv, SpawnStream
from oracle import cz_policy
from tests.replay import GOLDEN_DIR, load_golden


def policy_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "policy_*.npz")))


def oracle_env(cfg, layout, n):
    sp = cfg.get("spawn")
    kw = {} if not sp else dict(agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                                grace_period=sp["grace"], spawn_stream=SpawnStream(sp["seed"], n, 1))
    return OracleEnv(layout, cfg["recipes"], cfg["max_steps"], reward_scheme=cfg["reward_scheme"],
                     end_condition_all_dishes=cfg["end_all"], action_scheme=cfg.get("action_scheme", "scheme3"), **kw)


@pytest.mark.parametrize("path", policy_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_policy_oracle_matches_recorded_cook(path):
    g = load_golden(path)
    cfg = g["config"]
    names = cfg.get("policy_recipes", cfg["recipes"])
    seen = set()
    for n, layout in enumerate(g["layouts"]):
        env = oracle_env(cfg, layout, n)
        for t in range(int(g["length"][n])):
            got = cz_policy.heuristic_actions(env, names[:cfg["num_agents"]])
            want = g["policy"][n, t].tolist()
            assert got == want, f"{path} trace {n} step {t}: cook {want}, oracle {got}"
            seen.update(want)
            env.step(g["actions"][n, t])
    assert seen - {-1, 0}, "the traces must contain real decisions"


def test_policy_goldens_cover_crashes_and_all_moves():
    seen = set()
    for path in policy_files():
        seen.update(np.unique(load_golden(path)["policy"]).tolist())
    assert seen == {-1, 0, 1, 2, 3, 4}
