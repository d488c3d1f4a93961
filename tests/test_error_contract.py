"""The error contract: where the reference itself raises, the step is defined and flagged instead (include/cz_b200.h
CZ_ERR_*).  Each fixture was recorded from the unmodified reference up to the step at which it raised (`raised`,
`raised_type` in the .npz; `tests/golden/make_golden.py errors|error_bits`); here both oracles (CPU) and the CUDA kernels
(GPU, warp-per-environment and lane-per-environment) take that step: the matching bit must go up — and nowhere else —
and the three implementations must agree on the defined outputs.

Not reachable under the shipped classes and therefore not driven high here: CZ_ERR_CUTBOARD_NONE (Cutboard READY with a
first item that is not choppable: only ChopFood is ever accepted, world_objects.py:271-273) and CZ_ERR_REMOVE (scooping
an object that is not content of the faced static object: every dynamic object on a non-walkable cell is content of it).
"""
import json
import os

import numpy as np
import pytest

from oracle.cz_oracle import OracleEnv, SpawnStream
from oracle.cz_oracle_c import COracleEnv
from tests.replay import GOLDEN_DIR, ROOT, load_golden, bits, assert_state_equal, STATE_KEYS

ERR = {"CUTBOARD_NONE": 1, "REMOVE": 2, "SWITCH_LINK": 4, "SPAWN_LOC": 8, "TRUNC_DESPAWN": 16, "OBS_OVERFLOW": 32,
       "OFFGRID": 64, "BAD_ID": 128}
FIXTURES = {"c9_trunc_despawn": ("TRUNC_DESPAWN", "IndexError"), "c9_trunc_despawn_open4": ("TRUNC_DESPAWN", "IndexError"),
            "err_offgrid_primary": ("OFFGRID", "IndexError"), "err_offgrid_execute": ("OFFGRID", "IndexError"),
            "err_offgrid_special": (None, ""), "err_spawn_loc": ("SPAWN_LOC", "ValueError"),
            "err_switch_link": ("SWITCH_LINK", "AttributeError")}


def _golden(name):
    g = load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    g["raised_type"] = json.loads(str(g["raised_type"]))
    return g


def _oracle(cls, g, n):
    cfg, sp = g["config"], g["config"].get("spawn")
    kw = {} if not sp else dict(agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"],
                                grace_period=sp["grace"], spawn_stream=SpawnStream(sp["seed"], n, 1))
    return cls(g["layouts"][n], cfg["recipes"], cfg["max_steps"], reward_scheme=cfg["reward_scheme"],
               end_condition_all_dishes=cfg["end_all"], action_scheme=cfg.get("action_scheme", "scheme3"), **kw)


def _run_oracle(cls, g, n):
    """replay the recorded steps (already pinned by the golden replay tests), then the raising one"""
    env = _oracle(cls, g, n)
    A = g["config"]["num_agents"]
    for t in range(int(g["length"][n])):
        for i in range(A):
            if g["teleport"][n, t, i, 0] >= 0:
                env.teleport(i, *map(int, g["teleport"][n, t, i]))
        env.step(g["actions"][n, t])
    assert env.error == 0
    out = None
    if g["raised"][n] >= 0:
        rew, te, tu, rel = env.step(g["actions"][n, int(g["raised"][n])])
        out = (np.asarray(rew, np.float64), np.asarray(te, np.uint8), np.asarray(tu, np.uint8), np.asarray(rel, np.uint8),
               env.export_state(), np.stack([env.observe(i) for i in range(A)]))
    return env, out


@pytest.mark.parametrize("name", list(FIXTURES))
def test_oracles_flag_exactly_the_steps_where_the_reference_raised(name):
    g = _golden(name)
    bit, exc = FIXTURES[name]
    n_raised = 0
    for n in range(len(g["layouts"])):
        py, out_py = _run_oracle(OracleEnv, g, n)
        c, out_c = _run_oracle(COracleEnv, g, n)
        if g["raised"][n] < 0:
            assert py.error == 0 and c.error == 0, (name, n)
            continue
        n_raised += 1
        assert g["raised_type"][n] == exc
        assert py.error == ERR[bit] and c.error == ERR[bit], (name, n, py.error, c.error)
        for a, b in zip(out_py[:4], out_c[:4]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (name, n)
        assert_state_equal(out_py[4], out_c[4], f"{name} trace {n} raising step")
        assert np.array_equal(bits(out_py[5]), bits(out_c[5]))
        if bit == "TRUNC_DESPAWN":
            # the definition: every relevant agent is truncated, nobody stays active, the episode is over
            rew, te, tu, rel, st, _ = out_py
            assert np.array_equal(tu, rel) and rel.any() and not st["agents"][:, 4].any()
    assert (n_raised > 0) == (bit is not None)


def test_obs_overflow_is_flagged_when_a_bread_twin_has_no_slot():
    """meta file with two Bread slots, two Breads in the level: chopping one would append a third Bread and the
    reference's observation silently grows by five elements (cooking_env.py:371).  The oracles follow the reference
    (the twin exists) and flag the environment — the Python one when the vector is built, the C one at creation; the
    device has no slot for the twin, does not create it, and sets CZ_ERR_OBS_OVERFLOW at the chop.  Parity beyond
    that step is undefined by construction."""
    g = _golden("kat_coop_seed0")
    lay = dict(g["layouts"][0], meta=[[k, 2 if k == "Bread" else v] for k, v in g["layouts"][0]["meta"]])
    envs = [cls(lay, g["config"]["recipes"], 400, end_condition_all_dishes=True) for cls in (OracleEnv, COracleEnv)]
    for e in envs:
        for t in range(14, 17):
            for i in range(2):
                if g["teleport"][0, t, i, 0] >= 0:
                    e.teleport(i, *map(int, g["teleport"][0, t, i]))
            assert e.error == 0
            e.step(g["actions"][0, t])
            e.observe(0)
        assert e.error == ERR["OBS_OVERFLOW"]


@pytest.mark.reference
def test_reference_grows_its_observation_where_obs_overflow_is_flagged():
    from oracle.ref_harness import RefEnv
    g = _golden("kat_coop_seed0")
    ref = RefEnv(0, "coop_test", os.path.join(ROOT, "tests/golden/levels/meta_bread2.json"), 2, 400,
                 g["config"]["recipes"], end_condition_all_dishes=True)
    L0 = ref.observe_all().shape[1]
    for t in range(14, 17):
        for i in range(2):
            if g["teleport"][0, t, i, 0] >= 0:
                ref.teleport(i, *map(int, g["teleport"][0, t, i]))
        ref.step(g["actions"][0, t])
    assert L0 == 268 and ref.observe_all().shape[1] == L0 + 5


# ------------------------------------------------------------------------------------------------- GPU
def _device(g, n_envs, **kw):
    from cooking_zoo_b200 import BatchedCookingEnv
    cfg, sp = g["config"], g["config"].get("spawn")
    if sp:
        kw = dict(kw, agent_respawn_rate=sp["respawn"], agent_despawn_rate=sp["despawn"], grace_period=sp["grace"], seed=sp["seed"])
    return BatchedCookingEnv(n_envs, cfg["level"], cfg["meta_file"], cfg["num_agents"], cfg["max_steps"], cfg["recipes"],
                             end_condition_all_dishes=cfg["end_all"], reward_scheme=cfg["reward_scheme"],
                             action_scheme=cfg.get("action_scheme", "scheme3"), layouts=g["layouts"], **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["warp", "lane"])
@pytest.mark.parametrize("name", list(FIXTURES))
def test_device_flags_exactly_the_steps_where_the_reference_raised(name, kernel, monkeypatch):
    import torch
    if kernel == "lane":
        monkeypatch.setenv("CZ_WARP_MAX_ENVS", "0")
    g = _golden(name)
    bit, _ = FIXTURES[name]
    n, A = len(g["layouts"]), g["config"]["num_agents"]
    env = _device(g, n)
    env.reset(layout_ids=np.arange(n))
    T = int(g["length"].max())
    for t in range(T + 1):
        # environments past their recorded length keep receiving no-ops: they are not compared any more, except the
        # raising step itself, which every trace takes at t == raised
        act = np.zeros((n, A), np.uint8)
        for k in range(n):
            if t < g["length"][k] or t == g["raised"][k]:
                act[k] = g["actions"][k, t]
                for i in range(A):
                    if t < g["length"][k] and g["teleport"][k, t, i, 0] >= 0:
                        env.teleport(k, i, *map(int, g["teleport"][k, t, i]))
        before = env.error_flags.cpu().numpy().copy()
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act))
        flags = env.error_flags.cpu().numpy()
        for k in range(n):
            if t < g["length"][k]:
                assert flags[k] == 0, (name, k, t)
            elif t == g["raised"][k]:
                assert before[k] == 0 and flags[k] == ERR[bit], (name, k, t, flags[k])
                _, out = _run_oracle(OracleEnv, g, k)
                assert np.array_equal(bits(out[0]), bits(rew[k].cpu().numpy())), (name, k)
                assert np.array_equal(out[1], term[k].cpu().numpy()) and np.array_equal(out[2], trunc[k].cpu().numpy())
                assert_state_equal(out[4], env.export_state(env=k), f"{name} trace {k} raising step")
                assert np.array_equal(bits(out[5]), bits(obs[k].cpu().numpy()))
    if bit is None:
        assert int(env.error_flags.abs().sum()) == 0


@pytest.mark.gpu
def test_device_obs_overflow_and_bad_ids():
    import torch
    from cooking_zoo_b200 import BatchedCookingEnv
    g = _golden("kat_coop_seed0")
    env = BatchedCookingEnv(3, "coop_test", os.path.join(ROOT, "tests/golden/levels/meta_bread2.json"), 2, 400,
                            g["config"]["recipes"], end_condition_all_dishes=True, action_scheme="scheme3",
                            layouts=[dict(g["layouts"][0], meta=[[k, 2 if k == "Bread" else v] for k, v in g["layouts"][0]["meta"]])])
    env.reset(layout_ids=np.zeros(3, np.int32))
    for t in range(14, 17):
        for i in range(2):
            if g["teleport"][0, t, i, 0] >= 0:
                env.teleport(1, i, *map(int, g["teleport"][0, t, i]))        # only environment 1 plays the scenario
        assert int(env.error_flags.abs().sum()) == 0
        act = np.zeros((3, 2), np.uint8)
        act[1] = g["actions"][0, t]
        env.step(torch.from_numpy(act))
    assert env.error_flags.cpu().tolist() == [0, ERR["OBS_OVERFLOW"], 0]
    # ids outside the compiled tables passed as DEVICE tensors (host arrays are rejected by the Python wrapper):
    # the kernel uses id 0 and raises CZ_ERR_BAD_ID for that environment
    env.error_flags.zero_()
    lid = torch.tensor([0, 7, 0], dtype=torch.int32, device="cuda")
    obs = env.reset(layout_ids=lid).clone()
    assert env.error_flags.cpu().tolist() == [0, ERR["BAD_ID"], 0]
    assert torch.equal(obs[1].view(torch.int64), obs[0].view(torch.int64))
    env.error_flags.zero_()
    rid = torch.tensor([[0, 1], [0, 1], [9, 1]], dtype=torch.uint8, device="cuda")
    env.reset(layout_ids=torch.zeros(3, dtype=torch.int32, device="cuda"), recipe_ids=rid)
    assert env.error_flags.cpu().tolist() == [0, 0, ERR["BAD_ID"]]
