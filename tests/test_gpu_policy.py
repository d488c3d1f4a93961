"""GPU: the device policy (cz_policy_act) against the recorded decisions of the reference's
CookingAgent (tests/golden/policy_*.npz) and against the policy oracle on live batches."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.cz_oracle import OracleEnv, SpawnStream
from oracle import cz_policy
from tests.replay import GOLDEN_DIR, ROOT, load_golden
from tests.test_gpu_parity import _make

pytestmark = pytest.mark.gpu
BOOK = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana",
        "CucumberOnion", "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]


def policy_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "policy_*.npz")))


def _decode(actions, crashed, A):
    """device (actions, crashed bits) -> the recorder's convention: -1 where the cook raised"""
    a = actions.cpu().numpy().astype(np.int64)
    c = crashed.cpu().numpy()
    for i in range(A):
        bad = (c >> i) & 1 == 1
        assert (a[bad, i] == 0).all()
        a[bad, i] = -1
    return a


@pytest.mark.parametrize("path", policy_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_device_policy_replays_recorded_cook(path):
    g = load_golden(path)
    cfg = g["config"]
    n, A = len(g["layouts"]), cfg["num_agents"]
    names = cfg.get("policy_recipes", cfg["recipes"])[:A]
    env = _make(n, cfg, layouts=g["layouts"], recipe_pool=list(dict.fromkeys(list(cfg["recipes"]) + list(names))))
    env.reset(layout_ids=np.arange(n))
    explicit = names != cfg["recipes"][:A]
    for t in range(g["actions"].shape[1]):
        live = [k for k in range(n) if t < g["length"][k]]
        if not live:
            break
        got = _decode(*env.heuristic_actions(names if (explicit or t % 2) else None), A)
        for k in live:
            assert got[k].tolist() == g["policy"][k, t].tolist(), f"{path} trace {k} step {t}"
        env.step(torch.from_numpy(g["actions"][:, t].astype(np.uint8)))


def _live_lockstep(cfg, n_envs, check, steps, seed, pool, eps=0.15, per_env_cooks=False, **kw):
    A = cfg["num_agents"]
    env = _make(n_envs, cfg, layout_pool_size=48, layout_seed=seed, recipe_pool=pool, **kw)
    rng = np.random.default_rng(seed)
    lids = rng.integers(0, env.tables.num_layouts, size=n_envs).astype(np.int32)
    R = len(cfg["recipes"])
    rids = rng.integers(0, len(pool), size=(n_envs, R)).astype(np.uint8)
    cooks = rng.integers(0, len(pool), size=(n_envs, A)).astype(np.uint8) if per_env_cooks else rids[:, :A]
    env.reset(layout_ids=lids, recipe_ids=rids)
    names = env.tables.recipe_names
    picks = np.linspace(0, n_envs - 1, check).astype(int)
    sp = cfg.get("spawn")
    oracles = {}
    for k in picks:
        okw = {} if not sp else dict(agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                                     grace_period=sp["grace"], spawn_stream=SpawnStream(sp["seed"], int(k), 1))
        oracles[int(k)] = OracleEnv(env.tables.layouts[lids[k]], [names[r] for r in rids[k]], cfg["max_steps"],
                                    end_condition_all_dishes=cfg["end_all"],
                                    action_scheme=cfg.get("action_scheme", "scheme3"), **okw)
    cook_t = torch.from_numpy(cooks).cuda() if per_env_cooks else None
    seen = set()
    alive = set(oracles)
    for t in range(steps):
        got = _decode(*env.heuristic_actions(cook_t), A)
        for k in sorted(alive):
            want = cz_policy.heuristic_actions(oracles[k], [names[r] for r in cooks[k]])
            assert got[k].tolist() == want, f"env {k} step {t}: oracle {want}, device {got[k].tolist()}"
            seen.update(want)
        act = np.where(rng.random((n_envs, A)) < eps, rng.integers(0, 5, size=(n_envs, A)), np.maximum(got, 0)).astype(np.uint8)
        _, _, term, trunc, _ = env.step(torch.from_numpy(act))
        term, trunc = term.cpu().numpy(), trunc.cpu().numpy()
        for k in sorted(alive):
            _, te, tu, _ = oracles[k].step(act[k])
            assert [int(v) for v in te] == list(term[k]), f"env {k} step {t}"
            if any(te) or oracles[k].t >= cfg["max_steps"]:
                alive.discard(k)
        if not alive:
            break
    return seen


def test_device_policy_vs_oracle_coop_book():
    """2048 coop_test envs, per-env recipe pairs from the whole book, cooks follow their env's recipes"""
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=250,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    seen = _live_lockstep(cfg, 2048, 96, 250, 21, BOOK)
    assert {0, 1, 2, 3, 4} <= seen


def test_device_policy_vs_oracle_per_env_cooks_open4_spawn():
    """BASELINE config 5 shape: 4 agents, despawn / respawn, every cook with its own recipe (tensor argument)"""
    cfg = dict(level=os.path.join(ROOT, "tests/golden/levels/open4.json"),
               meta_file=os.path.join(ROOT, "tests/golden/levels/meta4.json"), num_agents=4, max_steps=200,
               recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"], end_all=True,
               reward_scheme=None, spawn={"respawn": 0.2, "despawn": 0.05, "grace": 3, "seed": 91})
    seen = _live_lockstep(cfg, 1024, 64, 200, 33, BOOK, per_env_cooks=True)
    assert {-1, 0, 1, 2, 3, 4} <= seen


def test_device_policy_vs_oracle_switch_and_optional_levels():
    for level, seed in (("switch_test", 41), ("coexistence_test", 43)):
        cfg = dict(level=level, meta_file="example", num_agents=2, max_steps=150,
                   recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
        _live_lockstep(cfg, 512, 48, 150, seed, BOOK[:3] + ["no_recipe"], per_env_cooks=True)


def test_device_cooks_finish_their_dishes_at_scale():
    """closed loop, no host in between: 16384 single-cook kitchens driven only by the device policy deliver
    TomatoLettuceSalad (terminated) well inside the step budget"""
    cfg = dict(level="coop_test", meta_file="example", num_agents=1, max_steps=400,
               recipes=["TomatoLettuceSalad"], end_all=False, reward_scheme=None)
    env = _make(16384, cfg, layout_pool_size=64)
    env.reset()
    done = torch.zeros(16384, dtype=torch.bool, device="cuda")
    for t in range(200):
        act, crashed = env.heuristic_actions()
        assert int(crashed.sum()) == 0
        _, _, term, _, _ = env.step(act)
        done |= term[:, 0].bool()
    assert float(done.float().mean()) > 0.95


def test_policy_argument_errors():
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=50,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    env = _make(8, cfg)
    env.reset()
    with pytest.raises(ValueError):
        env.heuristic_actions(["TomatoLettuceSalad"])                    # one name per agent
    with pytest.raises(ValueError):
        env.heuristic_actions(["TomatoLettuceSalad", "AppleWatermelon"])   # not in the compiled recipe pool
    with pytest.raises(ValueError):
        env.heuristic_actions(torch.zeros((8, 3), dtype=torch.uint8))
    # a book index outside the pool: the cook cannot exist -> reported as crashed, action 0
    act, crashed = env.heuristic_actions(torch.full((8, 2), 200, dtype=torch.uint8))
    assert crashed.tolist() == [3] * 8 and int(act.sum()) == 0


def test_cfg5_mixed_agent_counts_heuristic_streams_spawning_per_env_recipes():
    """BASELINE config 5 end to end on the device: 1-4 agents per environment, actions from the device cook
    (epsilon-mixed), despawn / respawn from the shared stream, a recipe assignment per environment; sampled
    environments in lockstep with the step oracle and the policy oracle."""
    from cooking_zoo_b200 import MixedAgentCookingEnv
    level = os.path.join(ROOT, "tests/golden/levels/open4.json")
    meta = os.path.join(ROOT, "tests/golden/levels/meta4.json")
    rng = np.random.default_rng(77)
    N, steps, seed = 1500, 120, 1234
    counts = rng.integers(1, 5, size=N)
    recipes = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
    env = MixedAgentCookingEnv(counts, level, meta, 10000, recipes, end_condition_all_dishes=True,
                               action_scheme="scheme3", recipe_pool=BOOK, layout_pool_size=40, layout_seed=3,
                               agent_respawn_rate=0.2, agent_despawn_rate=0.05, grace_period=3, seed=seed)
    assert sorted(env.groups) == [1, 2, 3, 4]
    lids = rng.integers(0, 40, size=N).astype(np.int32)
    rids = rng.integers(0, len(BOOK), size=(N, 4)).astype(np.uint8)
    obs = env.reset(layout_ids=lids, recipe_ids=rids)
    picks = rng.choice(N, size=60, replace=False)
    pos = {a: {int(g): j for j, g in enumerate(env.index[a].cpu().numpy())} for a in env.groups}
    oracles = {}
    for k in picks:
        a = int(counts[k])
        names = env.groups[a].tables.recipe_names
        oracles[int(k)] = OracleEnv(env.groups[a].tables.layouts[lids[k]], [names[r] for r in rids[k, :a]], 10000,
                                    end_condition_all_dishes=True, agent_respawn_rate=0.2, agent_despawn_rate=0.05,
                                    grace_period=3, spawn_stream=SpawnStream(seed, int(env.stream_ids[k]), 1))
    alive = set(oracles)
    for t in range(steps):
        act, crashed = env.heuristic_actions()
        act = act.cpu().numpy().astype(np.int64)
        crashed = crashed.cpu().numpy()
        for k in sorted(alive):
            a = int(counts[k])
            names = env.groups[a].tables.recipe_names
            want = cz_policy.heuristic_actions(oracles[k], [names[r] for r in rids[k, :a]])
            got = [(-1 if crashed[k] >> i & 1 else int(act[k, i])) for i in range(a)]
            assert got == want, f"env {k} ({a} agents) step {t}"
            assert (act[k, a:] == 0).all()
        mix = np.where(rng.random((N, 4)) < 0.1, rng.integers(0, 5, size=(N, 4)), act).astype(np.uint8)
        obs, rew, term, trunc = env.step(torch.from_numpy(mix))
        rew, term, trunc = rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        obs = {a: o.cpu().numpy() for a, o in obs.items()}
        for k in sorted(alive):
            a = int(counts[k])
            r, te, tu, _ = oracles[k].step(mix[k, :a])
            ctx = f"env {k} ({a} agents) step {t}"
            assert np.array_equal(np.asarray(r, np.float64).view(np.uint64), rew[k, :a].view(np.uint64)), ctx
            assert [int(v) for v in te] == list(term[k, :a]) and [int(v) for v in tu] == list(trunc[k, :a]), ctx
            want = np.stack([oracles[k].observe(i) for i in range(a)])
            assert np.array_equal(want.view(np.uint64), obs[a][pos[a][k]].view(np.uint64)), ctx
            if any(te):
                alive.discard(k)
    assert len(alive) > 20


@pytest.mark.parametrize("pipelined", [False, True])
def test_cook_steps_equals_the_step_by_step_closed_loop(pipelined):
    """MixedAgentCookingEnv.cook_steps(k) (one fork / join of the group streams around k steps) == k x cook_step() ==
    heuristic_actions() + step() through the global views: same states, rows, rewards and flags in every group"""
    from cooking_zoo_b200 import MixedAgentCookingEnv
    level = os.path.join(ROOT, "tests/golden/levels/open4.json")
    meta = os.path.join(ROOT, "tests/golden/levels/meta4.json")
    counts = (np.arange(3000) % 4) + 1
    recipes = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "AppleWatermelon"]

    def make(pipe):
        env = MixedAgentCookingEnv(counts, level, meta, 40, recipes, end_condition_all_dishes=True, action_scheme="scheme3",
                                   layout_pool_size=32, auto_reset=True, seed=11, agent_respawn_rate=0.2,
                                   agent_despawn_rate=0.05, grace_period=3, pipelined=pipe)
        env.reset()
        return env
    a, b, c = make(pipelined), make(pipelined), make(False)
    for _ in range(3):
        a.cook_steps(7)
        for _ in range(7):
            b.cook_step()
            act, _ = c.heuristic_actions()
            c.step(act)
        a.wait(); b.wait()
        torch.cuda.synchronize()
        for n in a.groups:
            ga, gb, gc = a.groups[n], b.groups[n], c.groups[n]
            for other in (gb, gc):
                assert torch.equal(ga.state, other.state), n
                assert torch.equal(ga.obs.view(torch.int64), other.obs.view(torch.int64)), n
                assert torch.equal(ga.reward.view(torch.int64), other.reward.view(torch.int64)), n
                assert torch.equal(ga.terminated, other.terminated) and torch.equal(ga.truncated, other.truncated), n
            assert int(ga.error_flags.abs().sum()) == 0


def test_pipelined_closed_loop_equals_in_place_closed_loop():
    """the cook only waits for the dynamics of the pipelined step (cz_pipeline_wait_state): same trajectories"""
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=60,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    n = 30000
    a = _make(n, cfg, auto_reset=True, seed=3, layout_pool_size=64)
    b = _make(n, cfg, auto_reset=True, seed=3, layout_pool_size=64, pipelined=True)
    a.reset(); b.reset()
    for t in range(90):
        aa, _ = a.heuristic_actions()
        ab, _ = b.heuristic_actions()
        assert torch.equal(aa, ab), t
        oa, ra, *_ = a.step(aa)
        ob, rb, *_ = b.step(ab)
        if t % 10 == 9:
            b.wait()
            assert torch.equal(oa.view(torch.int64), ob.view(torch.int64)) and torch.equal(ra, rb), t
    b.wait()
    assert torch.equal(a.state, b.state)


def test_strided_policy_grid_gives_the_same_decisions():
    """cz_policy_config: a grid of b blocks per SM that walks the batch in strides (background policy; measured and not
    used by default, profiles/r02_notes.md) decides exactly like the one-thread-per-environment launch"""
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=400,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    n = 70001
    a = _make(n, cfg, seed=4, layout_pool_size=64)
    b = _make(n, cfg, seed=4, layout_pool_size=64, background_policy=1)
    a.reset(); b.reset()
    for t in range(12):
        act_a, cr_a = a.heuristic_actions()
        act_b, cr_b = b.heuristic_actions()
        assert torch.equal(act_a, act_b) and torch.equal(cr_a, cr_b), t
        a.step(act_a); b.step(act_b)
    assert torch.equal(a.state, b.state)
