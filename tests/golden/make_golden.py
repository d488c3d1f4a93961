"""Generate golden traces from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference has no tests or golden vectors of its own, so these traces —
recorded by importing /root/reference behind oracle/refshim and driving
CookingEnvironment.accumulated_step / get_feature_vector directly — are the
pins for oracle/cz_oracle.py and for the CUDA path.  /root/reference does not exist on the
GPU box; the .npz files travel instead.

Each file holds `n` traces of one configuration:
    config   JSON: level, meta_file, num_agents, max_steps, recipes, end_all, reward_scheme
    layouts  JSON list: describe_layout() of every trace's initial world
    actions  i8 [n,T,A]      teleport i8 [n,T,A,2] (-1 = none; applied before the step)
    length   i32[n]          number of valid steps (trace stops at terminated/truncated)
    agents i16[n,T+1,A,6]  objs i16[n,T+1,D,9]  statics i16[n,T+1,S,4]  marks i32[n,T+1,R]
    reward f64[n,T,A]  term u8[n,T,A]  trunc u8[n,T,A]  rel u8[n,T,A]  obs f64[n,T+1,A,L]
index 0 along the T+1 axis is the state right after reset().
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import RefEnv  # noqa: E402
from oracle.cz_oracle import SpawnStream  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# Scripted single-agent solve of TomatoLettuceSalad on the coop_test layout of seed 3
# (SURVEY.md Appendix E16).  Moves are (dx,dy)->action: 1 left, 2 right, 3 down, 4 up.


def _path(p):
    """levels / meta files named by a repo-relative .json path are resolved against the repo root"""
    return os.path.join(ROOT, p) if p.endswith(".json") else p


def register_reference_recipes(custom):
    """cfg["custom_recipes"] = {name: (type, condition, [children])} -> the reference's own registry
    (recipe_drawer.register_recipe, :34-35).  Returns a callable that empties RECIPE_STORE again."""
    from oracle.ref_harness import load_reference
    load_reference()                       # puts /root/reference (behind oracle/refshim) on the path
    from cooking_zoo.cooking_book import recipe_drawer as rd
    from cooking_zoo.cooking_book.recipe import Recipe, RecipeNode
    from cooking_zoo.cooking_world import world_objects as wo
    from cooking_zoo.cooking_world.constants import ChopFoodStates, BlenderFoodStates
    conds = {"chopped": [("chop_state", ChopFoodStates.CHOPPED)], "mashed": [("blend_state", BlenderFoodStates.MASHED)],
             None: None}

    def build(t):
        typ, cond, kids = t
        return RecipeNode(root_type=getattr(wo, typ), id_num=rd.get_next_id(), name=typ, conditions=conds[cond],
                          contains=[build(k) for k in kids])
    roots = {name: build(t) for name, t in custom.items()}
    for name, root in roots.items():
        rd.register_recipe(Recipe(root, rd.NUM_GOALS), name)
    # cooking_env copies NUM_GOALS by value when it is imported (cooking_env.py:7), so a user has to register before
    # importing the environment; the harness imported it already, so the copy is refreshed to what that order gives
    import cooking_zoo.environment.cooking_env as ce
    stale = ce.NUM_GOALS
    ce.NUM_GOALS = rd.NUM_GOALS

    def undo():
        rd.RECIPE_STORE.clear()
        ce.NUM_GOALS = stale
    return undo


def record(cfg, seeds, T, policy, teleports=None):
    if cfg.get("custom_recipes"):
        undo = register_reference_recipes(cfg["custom_recipes"])
        try:
            return record({k: v for k, v in cfg.items() if k != "custom_recipes"}, seeds, T, policy, teleports)
        finally:
            undo()
    traces = []
    for n_trace, seed in enumerate(seeds):
        spawn = cfg.get("spawn")      # {"respawn", "despawn", "grace", "seed"}: trace n is environment n of the stream
        env = RefEnv(seed, _path(cfg["level"]), _path(cfg["meta_file"]), cfg["num_agents"], cfg["max_steps"],
                     cfg["recipes"], end_condition_all_dishes=cfg["end_all"],
                     reward_scheme=cfg.get("reward_scheme"), action_scheme=cfg.get("action_scheme", "scheme3"),
                     **({} if not spawn else dict(agent_respawn_rate=spawn["respawn"], agent_despawn_rate=spawn["despawn"],
                                                 grace_period=spawn["grace"],
                                                 spawn_stream=SpawnStream(spawn["seed"], n_trace, 1))))
        if hasattr(policy, "bind"):
            policy.bind(env, cfg)
        A = cfg["num_agents"]
        rng = np.random.default_rng(1000 + seed)
        tr = {"layout": env.layout(), "actions": np.zeros((T, A), np.int8),
              "teleport": -np.ones((T, A, 2), np.int8)}
        if hasattr(policy, "raw"):
            tr["policy"] = np.zeros((T, A), np.int8)
        st = env.export_state()
        keys = ("agents", "objs", "statics", "marks")
        hist = {k: [st[k]] for k in keys}
        obs = [env.observe_all()]
        rew, term, trunc, rel = [], [], [], []
        length = T
        prev = np.zeros(A, np.int64)
        for t in range(T):
            if teleports is not None and t in teleports:
                for i, (x, y) in teleports[t].items():
                    env.teleport(i, x, y)
                    tr["teleport"][t, i] = (x, y)
            act = policy(rng, t, A, prev)
            prev = act
            tr["actions"][t] = act
            if "policy" in tr:
                tr["policy"][t] = policy.raw
            try:
                r, te, tu, re_ = env.step(act)
            except (IndexError, ValueError, AttributeError) as ex:
                tr["raised_type"] = type(ex).__name__
                # truncation with a despawned agent raises in the reference
                # (cooking_env.py:337/348); the trace ends before that step and remembers it: the action of
                # step `raised` stays in `actions`, so a replay can take the step and check the error contract
                length = t
                tr["raised"] = t
                break
            rew.append(r); term.append(te); trunc.append(tu); rel.append(re_)
            st = env.export_state()
            for k in keys:
                hist[k].append(st[k])
            obs.append(env.observe_all())
            # the episode ends on termination or when the clock runs out; a despawning agent is
            # truncated individually (cooking_env.py:346-349) and the environment carries on
            if te.any() or env.env.t >= cfg["max_steps"]:
                length = t + 1
                break
        tr["length"] = length
        for k in keys:
            tr[k] = np.stack(hist[k])
        tr["obs"] = np.stack(obs)
        tr["reward"] = np.stack(rew); tr["term"] = np.stack(term)
        tr["trunc"] = np.stack(trunc); tr["rel"] = np.stack(rel)
        traces.append(tr)
    return traces


def save(name, cfg, traces, T):
    n = len(traces)

    def pad(key, axis_len):
        first = traces[0][key]
        out = np.zeros((n, axis_len) + first.shape[1:], first.dtype)
        for i, tr in enumerate(traces):
            out[i, :tr[key].shape[0]] = tr[key]
        return out

    arrays = {
        "config": np.array(json.dumps(cfg)),
        "layouts": np.array(json.dumps([tr["layout"] for tr in traces])),
        "actions": np.stack([tr["actions"] for tr in traces]),
        "teleport": np.stack([tr["teleport"] for tr in traces]),
        "length": np.array([tr["length"] for tr in traces], np.int32),
        # step at which the reference itself raised IndexError (Appendix C-9), -1: it did not
        "raised": np.array([tr.get("raised", -1) for tr in traces], np.int32),
        "raised_type": np.array(json.dumps([tr.get("raised_type", "") for tr in traces])),
    }
    if "policy" in traces[0]:
        # raw CookingAgent.step outputs on the state BEFORE step t (-1: the reference agent raised)
        arrays["policy"] = np.stack([tr["policy"] for tr in traces])
    for k in ("agents", "objs", "statics", "marks", "obs"):
        arrays[k] = pad(k, T + 1)
    for k in ("reward", "term", "trunc", "rel"):
        arrays[k] = pad(k, T)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {n} traces, {os.path.getsize(path) / 1024:.0f} KiB")


def uniform(rng, t, A, prev):
    return rng.integers(0, 5, size=A)


def sticky(rng, t, A, prev):
    """repeat the previous action w.p. 0.5: more bumping into appliances, deeper states"""
    new = rng.integers(0, 5, size=A)
    keep = rng.random(A) < 0.5
    return np.where(keep & (t > 0), prev, new)


class Heuristic:
    """The reference's own scripted cook (cooking_agents/cooking_agent.py) on the live object graph,
    with an epsilon of uniform actions: the only streams here that finish recipes."""

    def __init__(self, eps):
        self.eps = eps
        self.raw = None     # the cooks' own decisions of the last call, before the epsilon mix

    def bind(self, env, cfg):
        from cooking_zoo.cooking_agents.cooking_agent import CookingAgent
        self.env = env
        names = cfg.get("policy_recipes", cfg["recipes"])     # the cooks' own recipes (default: the environment's)
        self.cooks = [CookingAgent(names[i], f"agent-{i + 1}") for i in range(cfg["num_agents"])]

    def __call__(self, rng, t, A, prev):
        from collections import defaultdict
        sym = defaultdict(list)
        sym.update(self.env.env.world.world_objects)
        sym["Agent"] = self.env.env.world.agents
        raw = []
        for c in self.cooks:
            try:
                raw.append(int(c.step(sym)))
            except (IndexError, AttributeError, TypeError):
                raw.append(-1)           # the scripted cook itself raises on this world (oracle/cz_policy.py)
        self.raw = np.array(raw)
        act = np.maximum(self.raw, 0)
        return np.where(rng.random(A) < self.eps, rng.integers(0, 5, size=A), act)


def scheme1_mix(rng, t, A, prev):
    """scheme1's eight actions, interaction-heavy, sticky"""
    new = rng.choice([0, 1, 2, 3, 4, 5, 5, 5, 6, 7, 7], size=A)
    keep = rng.random(A) < 0.3
    return np.where(keep & (t > 0), prev, new)


def scripted(seq):
    def pol(rng, t, A, prev):
        return np.array(seq[t] if t < len(seq) else [0] * A)
    return pol


def main_policy():
    """SURVEY §8 f3 / BASELINE config 5: traces that also hold the scripted cook's raw decisions"""
    base = {"level": "coop_test", "meta_file": "example", "max_steps": 400, "reward_scheme": None}
    book = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana",
            "CucumberOnion", "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]
    cfg = dict(base, num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, max_steps=300)
    save("policy_cfg2", cfg, record(cfg, range(1100, 1108), 300, Heuristic(0.1)), 300)
    for k in range(0, 8, 2):
        c = dict(cfg, recipes=book[k:k + 2])
        save(f"policy_book_{k}", c, record(c, range(1110 + k, 1113 + k), 300, Heuristic(0.15)), 300)
    # the cooks follow other recipes than the environment scores
    c = dict(cfg, recipes=["TomatoSalad", "no_recipe"], policy_recipes=["CarrotBanana", "TomatoLettuceSalad"])
    save("policy_other", c, record(c, range(1130, 1134), 300, Heuristic(0.1)), 300)
    # config 5: 1..4 agents in the open kitchen, despawn / respawn on, heuristic streams
    for a in (1, 2, 3, 4):
        c = dict(base, level="tests/golden/levels/open4.json", meta_file="tests/golden/levels/meta4.json",
                 num_agents=a, recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "AppleWatermelon"][:a],
                 end_all=True, max_steps=10000,
                 spawn={"respawn": 0.2, "despawn": 0.05 if a > 1 else 0.0, "grace": 3, "seed": 500 + a})
        save(f"policy_cfg5_a{a}", c, record(c, range(1140 + 4 * a, 1144 + 4 * a), 200, Heuristic(0.1)), 200)
    # scheme1 environment, same cook (it only ever walks)
    c = dict(cfg, action_scheme="scheme1")
    save("policy_scheme1", c, record(c, range(1170, 1174), 300, Heuristic(0.1)), 300)
    # OPTIONAL objects: worlds that lack an ingredient make the cook raise
    c = dict(base, level="coexistence_test", num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"],
             end_all=True, max_steps=250)
    save("policy_coexistence", c, record(c, range(1180, 1190), 250, Heuristic(0.1)), 250)
    c = dict(base, level="switch_test", num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"],
             end_all=True, max_steps=300)
    save("policy_switch", c, record(c, range(1190, 1194), 300, Heuristic(0.1)), 300)


def main_custom():
    """recipes registered through the reference's register_recipe hook (cooking_env.py:100-105): a root with two
    children that each have their own child, and a Counter root — shapes the book does not have.  The cooks follow book
    recipes with the same ingredients (CookingAgent reads RECIPES, not the store: base_agent.py:36)."""
    base = {"level": "coop_test", "meta_file": "example", "max_steps": 300, "num_agents": 2, "end_all": True}
    custom = {"TwoPlates": ("Deliversquare", None, [("Plate", None, [("Tomato", "chopped", [])]),
                                                    ("Plate", None, [("Lettuce", "chopped", [])])]),
              "CounterSalad": ("Counter", None, [("Plate", None, [("Tomato", "chopped", []), ("Lettuce", "chopped", [])])]),
              "BareCarrot": ("Carrot", "mashed", [])}
    cfg = dict(base, recipes=["TwoPlates", "CounterSalad"], custom_recipes=custom,
               policy_recipes=["TomatoLettuceSalad", "TomatoLettuceSalad"],
               reward_scheme={"recipe_reward": 20, "max_time_penalty": -5, "recipe_penalty": -40, "recipe_node_reward": 1.5})
    save("custom_recipes", cfg, record(cfg, range(1300, 1308), 300, Heuristic(0.15)), 300)
    cfg2 = dict(cfg, recipes=["BareCarrot", "TwoPlates"], end_all=False, policy_recipes=["MashedCarrotBanana", "TomatoSalad"])
    save("custom_recipes_any", cfg2, record(cfg2, range(1310, 1314), 300, Heuristic(0.15)), 300)


def main_errors():
    """runs that reach max_steps with despawn / respawn enabled.  Where an agent with an
    index >= len(env.agents) is relevant on the time-up step the reference raises IndexError (cooking_env.py:337
    sizes active_agents by the LIVE agent count, :348 indexes it by world agent index); the trace stores the step
    (`raised`).  Traces that time out with every relevant agent below that bound do not raise and are ordinary."""
    base = {"level": "coop_test", "meta_file": "example", "reward_scheme": None}
    cfg = dict(base, num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, max_steps=30,
               spawn={"respawn": 0.25, "despawn": 0.12, "grace": 2, "seed": 909})
    save("c9_trunc_despawn", cfg, record(cfg, range(1400, 1424), 30, sticky), 30)
    cfg4 = dict(base, level="tests/golden/levels/open4.json", meta_file="tests/golden/levels/meta4.json", num_agents=4,
                recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"], end_all=True, max_steps=24,
                spawn={"respawn": 0.2, "despawn": 0.15, "grace": 1, "seed": 910})
    save("c9_trunc_despawn_open4", cfg4, record(cfg4, range(1430, 1446), 24, uniform), 24)


def main_error_bits():
    """One directed trace per reachable CZ_ERR_* bit whose trigger makes the reference itself raise: the trace ends at the
    raising step (`raised`, `raised_type`).  tests/test_error_contract.py replays them and takes that step."""
    lv = "tests/golden/levels/"
    # CZ_ERR_OFFGRID: scheme1 interaction while facing a cell off the grid: get_objects_at(...)[0] IndexError
    # (cooking_world.py:119 / :143 / :160).  Agent 0 spawns on (0,0) facing left (orientation 1).
    cfg = {"level": lv + "edge_floor.json", "meta_file": "example", "max_steps": 50, "reward_scheme": None, "num_agents": 2,
           "recipes": ["TomatoSalad", "no_recipe"], "end_all": False, "action_scheme": "scheme1"}
    for name, first in (("err_offgrid_primary", 5), ("err_offgrid_special", 6), ("err_offgrid_execute", 7)):
        pol, tele = _script([({}, [3, 0]), ({}, [4, 4]), ({}, [1, 0]), ({}, [first, 0])], 2)
        save(name, cfg, record(cfg, [0], 4, pol, tele), 4)
    # CZ_ERR_SPAWN_LOC: agent 0 respawns while agent 1 stands on its only spawn cell: generate_location gives up after
    # 1001 draws (ValueError, parsing.py:154-167).  despawn / respawn rates 1.0, no grace period.
    cfg = {"level": lv + "one_cell_spawn.json", "meta_file": "example", "max_steps": 50, "reward_scheme": None,
           "num_agents": 2, "recipes": ["TomatoSalad", "no_recipe"], "end_all": False,
           "spawn": {"respawn": 1.0, "despawn": 1.0, "grace": 0, "seed": 5}}
    pol, tele = _script([({}, [3, 1]), ({}, [0, 0]), ({}, [0, 0])], 2)
    save("err_spawn_loc", cfg, record(cfg, [0], 3, pol, tele), 3)
    # CZ_ERR_SWITCH_LINK: two Switches in one level link to each other (level ATTRIBUTES are never applied, Appendix
    # C-6, so every LinkedObject shares group None): pressing one calls switch_state() on the other Switch, which
    # has no such method (AttributeError, world_objects.py:165-169)
    cfg = {"level": lv + "two_switch.json", "meta_file": "example", "max_steps": 50, "reward_scheme": None,
           "num_agents": 1, "recipes": ["TomatoSalad"], "end_all": False}
    pol, tele = _script([({}, [3]), ({}, [2])], 1)
    save("err_switch_link", cfg, record(cfg, [0], 2, pol, tele), 2)


def main_aec():
    """The PettingZoo AEC surface of the reference (cooking_env.py:26-43, 215-241: agent_selection, last(), per-agent
    step) recorded call by call: for every step() call the selected agent, what last() returned before it (observation,
    cumulative reward, terminated, truncated, info scalars) and the action passed.  Quirks this pins: the loop variable
    of :228 shadows `agent`, so the LAST agent's cumulative reward is zeroed on every call and the others' are never
    reset.  One file, two traces (uniform actions; the reference's own cooks, which finish a recipe)."""
    import random as _random
    from oracle.ref_harness import load_reference
    from oracle import ref_dump
    ce = load_reference()
    cfg = {"level": "coop_test", "meta_file": "example", "max_steps": 60, "reward_scheme": None, "num_agents": 2,
           "recipes": ["TomatoLettuceSalad", "CarrotBanana"], "end_all": False, "action_scheme": "scheme3"}
    traces = []
    for seed, pol in ((1500, uniform), (1501, Heuristic(0.05))):
        _random.seed(seed)
        np.random.seed(seed)
        env = ce.CookingEnvironment(level=cfg["level"], meta_file=cfg["meta_file"], num_agents=2, max_steps=cfg["max_steps"],
                                    recipes=cfg["recipes"], obs_spaces=["feature_vector"] * 2,
                                    end_condition_all_dishes=False, action_scheme="scheme3")
        env.reset()

        class _Shim:          # Heuristic.bind expects the RefEnv wrapper
            pass
        shim = _Shim()
        shim.env = env
        if hasattr(pol, "bind"):
            pol.bind(shim, cfg)
        rng = np.random.default_rng(seed)
        layout = ref_dump.describe_layout(env)
        calls = {"agent": [], "obs": [], "cum": [], "term": [], "trunc": [], "info_t": [], "info_done": [], "action": []}
        prev = np.zeros(2, np.int64)
        t = 0
        while True:
            act = pol(rng, t, 2, prev)
            prev = act
            t += 1
            finished = False
            for _ in range(len(env.agents)):
                name = env.agent_selection
                i = env.possible_agents.index(name)
                obs, cum, term, trunc, info = env.last()
                calls["agent"].append(i); calls["obs"].append(np.asarray(obs, np.float64)); calls["cum"].append(float(cum))
                calls["term"].append(int(term)); calls["trunc"].append(int(trunc))
                calls["info_t"].append(int(info.get("t", -1))); calls["info_done"].append(int(info.get("recipe_done", -1)))
                if term or trunc:
                    finished = True      # quirk C-8: the dead-agent loop cannot make progress; the trace ends here
                    calls["action"].append(-1)
                    break
                calls["action"].append(int(act[i]))
                env.step(int(act[i]))
            if finished:
                break
        traces.append((layout, {k: np.asarray(v) for k, v in calls.items()}))
    n = max(len(c["agent"]) for _, c in traces)
    out = {"config": np.array(json.dumps(cfg)), "layouts": np.array(json.dumps([l for l, _ in traces])),
           "n_calls": np.array([len(c["agent"]) for _, c in traces], np.int32)}
    for k in traces[0][1]:
        first = traces[0][1][k]
        arr = np.zeros((len(traces), n) + first.shape[1:], first.dtype)
        for j, (_, c) in enumerate(traces):
            arr[j, :len(c[k])] = c[k]
        out[k] = arr
    path = os.path.join(OUT, "aec", "aec_cfg2.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("aec_cfg2:", out["n_calls"].tolist(), "calls", os.path.getsize(path) // 1024, "KiB")


def _script(steps, A):
    """steps: list of (teleports {agent: (x, y)}, actions [A]) -> (scripted policy, teleports by step)"""
    tele = {t: tp for t, (tp, _) in enumerate(steps) if tp}
    return scripted([list(a) for _, a in steps]), tele


def main_kat():
    """directed known-answer scenarios, recorded from the reference with Agent.move_to teleports
    (layout of random.seed(0)).  tests/test_kat.py states what each step must show; the usual replays then hold the
    oracle, the C oracle and the CUDA kernels to every recorded array."""
    base = {"level": "coop_test", "meta_file": "example", "max_steps": 400, "reward_scheme": None}
    cfg2 = dict(base, num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True)
    steps = [
        ({0: (1, 2), 1: (1, 4)}, [3, 4]),   # t0  E1: both target (1,3): both cancelled, orientations still change
        ({0: (1, 2), 1: (1, 3)}, [3, 4]),   # t1  E2: swap
        ({0: (1, 2), 1: (1, 1)}, [4, 4]),   # t2  E3: a1 bumps the counter, a0 targets a1's cell: cancelled
        ({0: (2, 2), 1: (5, 1)}, [0, 2]),   # t3  E5: a1 grabs the Banana from counter (6,1)
        ({1: (4, 5)}, [0, 3]),              # t4  E6: Banana into the Blender (4,6): READY, toggle off, still FRESH
        ({}, [0, 3]),                       # t5  E7: execute: mashed in the same step, toggle off again, NOT_USABLE
        ({}, [0, 3]),                       # t6  E8: a1 holds the mashed Banana
        ({1: (5, 2)}, [0, 2]),              # t7  E9: counter (6,2) holds a Watermelon: nothing happens
        ({1: (4, 1)}, [0, 4]),              # t8  E10: Banana straight onto the Deliversquare (4,0)
        ({}, [0, 4]),                       # t9  E11: a Deliversquare never releases
        ({1: (5, 3)}, [0, 2]),              # t10      a1 grabs the Plate from counter (6,3)
        ({1: (4, 1)}, [0, 4]),              # t11 E12: the Banana is scooped onto the held Plate
        ({1: (5, 2)}, [0, 2]),              # t12 E13: Plate[Banana] vs fresh Watermelon: branch 2 rejects, no fall-through
        ({1: (4, 1)}, [0, 4]),              # t13 E14: Plate placed on the Deliversquare; CarrotBanana not complete
        ({0: (1, 3)}, [1, 0]),              # t14 E17: a0 grabs the Bread from counter (0,3)
        ({0: (1, 1)}, [1, 0]),              # t15      onto the Cutboard (0,1): READY
        ({}, [1, 0]),                       # t16      chop: a second, chopped Bread appears on the board
        ({}, [1, 0]),                       # t17      grab takes the new Bread (free), the old one stays
        ({0: (1, 3)}, [1, 0]),              # t18      put it on the empty counter (0,3)
        ({0: (1, 1)}, [1, 0]),              # t19      grab the first Bread
    ]
    pol, tele = _script(steps, 2)
    save("kat_coop_seed0", cfg2, record(cfg2, [0], len(steps), pol, tele), len(steps))
    # E4: three agents in the open kitchen: the third ends on the first one's cell
    cfg3 = dict(base, level="tests/golden/levels/open4.json", meta_file="tests/golden/levels/meta4.json", num_agents=3,
                recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad"], end_all=True)
    pol, tele = _script([({0: (1, 2), 1: (1, 4), 2: (2, 2)}, [3, 4, 1]), ({}, [0, 0, 0])], 3)
    save("kat_open4_three", cfg3, record(cfg3, [0], 2, pol, tele), 2)
    # E15: switch_test, one agent steps on the Switch (4,3) and then stands still: it toggles every step
    cfgs = dict(base, level="switch_test", num_agents=1, recipes=["TomatoLettuceSalad"], end_all=False)
    pol, tele = _script([({0: (3, 3)}, [2]), ({}, [0]), ({}, [0]), ({}, [1]), ({}, [2])], 1)
    save("kat_switch", cfgs, record(cfgs, [0], 5, pol, tele), 5)


def main():
    if sys.argv[1:] == ["errors"]:
        return main_errors()
    if sys.argv[1:] == ["kat"]:
        return main_kat()
    if sys.argv[1:] == ["error_bits"]:
        return main_error_bits()
    if sys.argv[1:] == ["aec"]:
        return main_aec()
    if sys.argv[1:] == ["policy"]:
        return main_policy()
    if sys.argv[1:] == ["custom"]:
        return main_custom()
    base = {"level": "coop_test", "meta_file": "example", "max_steps": 400, "reward_scheme": None}
    # BASELINE config 1: single agent, TomatoLettuceSalad
    cfg1 = dict(base, num_agents=1, recipes=["TomatoLettuceSalad"], end_all=False)
    save("cfg1_uniform", cfg1, record(cfg1, range(4), 400, uniform), 400)
    # BASELINE config 2: two agents, two recipes, all dishes
    cfg2 = dict(base, num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True)
    save("cfg2_uniform", cfg2, record(cfg2, range(12), 400, uniform), 400)
    save("cfg2_sticky", cfg2, record(cfg2, range(100, 112), 400, sticky), 400)
    # short max_steps: truncation path
    cfg3 = dict(cfg2, max_steps=25)
    save("cfg2_trunc25", cfg3, record(cfg3, range(200, 204), 25, sticky), 25)
    # non-default reward scheme with float terms and node rewards
    cfg4 = dict(cfg2, reward_scheme={"recipe_reward": 12.5, "max_time_penalty": -3.3,
                                     "recipe_penalty": -7.25, "recipe_node_reward": 1.1})
    save("cfg2_rewardscheme", cfg4, record(cfg4, range(300, 304), 400, sticky), 400)
    # every recipe of the book once (BASELINE config 4 draws per-env recipe pairs from it)
    book = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana",
            "CucumberOnion", "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]
    for k in range(0, 8, 2):
        cfg = dict(base, num_agents=2, recipes=book[k:k + 2], end_all=False, max_steps=200)
        save(f"book_{k}", cfg, record(cfg, range(400 + k, 402 + k), 200, sticky), 200)
    # heuristic cooks: recipes get completed -> bonus reward, termination (any / all dishes)
    cfgh = dict(base, num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=False, max_steps=200)
    save("heuristic_any", cfgh, record(cfgh, range(500, 506), 200, Heuristic(0.1)), 200)
    cfgh2 = dict(base, num_agents=2, recipes=["TomatoSalad", "AppleWatermelon"], end_all=True, max_steps=300)
    save("heuristic_all", cfgh2, record(cfgh2, range(510, 516), 300, Heuristic(0.05)), 300)
    cfgh1 = dict(base, num_agents=1, recipes=["TomatoLettuceSalad"], end_all=False, max_steps=200)
    save("heuristic_cfg1", cfgh1, record(cfgh1, range(520, 524), 200, Heuristic(0.0)), 200)
    # node rewards + penalty: a completed dish picked up again (recipe_undone)
    cfgp = dict(cfgh2, end_all=True, reward_scheme={"recipe_reward": 20, "max_time_penalty": -5,
                                                   "recipe_penalty": -40, "recipe_node_reward": 2})
    save("heuristic_nodes", cfgp, record(cfgp, range(530, 534), 300, Heuristic(0.15)), 300)
    # Switch / Block level
    cfgs = dict(base, level="switch_test", num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"],
                end_all=True, max_steps=300)
    save("switch_uniform", cfgs, record(cfgs, range(600, 606), 300, sticky), 300)
    # open 4-agent kitchen (custom level + meta by path): agent-agent collisions, three/four agents
    cfg4a = dict(base, level="tests/golden/levels/open4.json", meta_file="tests/golden/levels/meta4.json",
                 num_agents=4, recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"],
                 end_all=True, max_steps=250)
    save("open4_agents4", cfg4a, record(cfg4a, range(700, 706), 250, sticky), 250)
    cfg3a = dict(cfg4a, num_agents=3, recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad"])
    save("open4_agents3", cfg3a, record(cfg3a, range(710, 714), 250, uniform), 250)
    # tiny 4-agent room: crowded, constant collisions; exercises the NA=3/4 specialised kernels
    cfgt = dict(base, level="tests/golden/levels/tiny4.json", meta_file="tests/golden/levels/meta4.json",
                num_agents=4, recipes=["TomatoSalad", "TomatoSalad", "no_recipe", "no_recipe"], end_all=False,
                max_steps=150)
    save("tiny4_agents4", cfgt, record(cfgt, range(720, 726), 150, uniform), 150)
    cfgt3 = dict(cfgt, num_agents=3, recipes=["TomatoSalad", "no_recipe", "no_recipe"])
    save("tiny4_agents3", cfgt3, record(cfgt3, range(730, 734), 150, sticky), 150)
    # scheme1 (the constructor default, cooking_env.py:27): explicit primary / pick-up-special / execute actions
    cfgs1 = dict(cfg2, action_scheme="scheme1", max_steps=300)
    save("scheme1_cfg2", cfgs1, record(cfgs1, range(900, 910), 300, scheme1_mix), 300)
    cfgs1b = dict(cfg1, action_scheme="scheme1", max_steps=300)
    save("scheme1_cfg1", cfgs1b, record(cfgs1b, range(910, 914), 300, scheme1_mix), 300)
    cfgs1c = dict(cfgs, action_scheme="scheme1")
    save("scheme1_switch", cfgs1c, record(cfgs1c, range(920, 924), 300, scheme1_mix), 300)
    cfgs1d = dict(cfg4a, action_scheme="scheme1")
    save("scheme1_open4", cfgs1d, record(cfgs1d, range(930, 934), 250, scheme1_mix), 250)
    # coexistence_test: OPTIONAL objects (parsing.py:28-33,87-92) -> per-layout object sets and static variants
    cfgco = dict(base, level="coexistence_test", num_agents=2, recipes=["TomatoLettuceSalad", "CarrotBanana"],
                 end_all=True, max_steps=250)
    save("coexistence_sticky", cfgco, record(cfgco, range(940, 952), 250, sticky), 250)
    cfgco1 = dict(cfgco, action_scheme="scheme1")
    save("coexistence_scheme1", cfgco1, record(cfgco1, range(960, 966), 250, scheme1_mix), 250)
    # agent despawn / respawn (BASELINE config 5): randomness from the shared stream
    cfgsp = dict(cfg2, max_steps=10000, spawn={"respawn": 0.3, "despawn": 0.1, "grace": 2, "seed": 4242})
    save("spawn_cfg2", cfgsp, record(cfgsp, range(800, 808), 300, sticky), 300)
    cfgsp4 = dict(cfg4a, max_steps=10000, spawn={"respawn": 0.2, "despawn": 0.15, "grace": 3, "seed": 77})
    save("spawn_open4", cfgsp4, record(cfgsp4, range(810, 816), 200, uniform), 200)
    main_policy()
    main_custom()
    main_errors()
    main_kat()
    main_error_bits()
    main_aec()


if __name__ == "__main__":
    main()
