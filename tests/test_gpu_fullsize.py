"""GPU, BASELINE full size (131072 two-agent environments, the per-GPU shard of config 4): properties that
do not need the oracle to run at that size, plus a sampled lockstep against it."""
import numpy as np
import pytest
import torch

from oracle.cz_oracle import OracleEnv
from tests.replay import assert_obs_equal, bits

pytestmark = pytest.mark.gpu
N = 131072
BOOK = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana", "CucumberOnion",
        "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]


def _env(n=N, **kw):
    from cooking_zoo_b200 import BatchedCookingEnv
    return BatchedCookingEnv(n, "coop_test", "example", 2, 400, BOOK[1:3], end_condition_all_dishes=True, action_scheme="scheme3",
                             recipe_pool=BOOK, layout_pool_size=400, layout_seed=0, **kw)


def _actions(steps, n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, 5, (steps, n, 2), generator=g, dtype=torch.uint8).cuda()


def _recipe_ids(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, len(BOOK), (n, 2), generator=g, dtype=torch.uint8)


def _check_invariants(env):
    """Structural invariants of the packed state (include/cz_b200.h), evaluated with torch on the device:
    every object sits in exactly one container, held objects agree with their holder, content positions of
    a container are 0..n-1, exactly the last item of a container is `free`, positions are inside the level."""
    t = env.tables
    D, A = t.num_dyn_slots, t.num_agents
    st = env.state.to(torch.int64) & 0xFFFFFFFF
    o, ag = st[:D], st[D:D + A]
    pres = (o >> 6) & 1
    x, y = o & 7, (o >> 3) & 7
    ck, cid, pos, free = (o >> 10) & 3, (o >> 12) & 31, (o >> 17) & 63, (o >> 9) & 1
    assert bool(((x < t.width) & (y < t.height))[pres == 1].all())
    assert bool((ck[pres == 1] <= 2).all())
    # held objects: the named agent holds exactly this slot and stands on the same cell
    has, hold = (ag >> 9) & 1, (ag >> 10) & 31
    slots = torch.arange(D, device=st.device)[:, None]
    for i in range(A):
        held_by_i = (pres == 1) & (ck == 0) & (cid == i)
        assert bool((held_by_i.sum(0) == has[i]).all())
        assert bool((held_by_i == ((slots == hold[i][None, :]) & (has[i] == 1)[None, :])).all())
        same_cell = ((o & 63) == (ag[i] & 63)[None, :])
        assert bool(same_cell[held_by_i].all())
    # plate content: the container is a present plate, shares its cell, positions are a permutation
    tf = torch.as_tensor(t.type_flags[t.slot_type[:D]].astype(np.int64), device=st.device)
    in_plate = (pres == 1) & (ck == 2)
    plate_idx = torch.where(in_plate, cid, torch.zeros_like(cid))
    plate_rec = torch.gather(o, 0, plate_idx)
    assert bool(((tf[plate_idx] & 1) == 1)[in_plate].all()) and bool((((plate_rec >> 6) & 1) == 1)[in_plate].all())
    assert bool(((plate_rec & 63) == (o & 63))[in_plate].all())
    for p in torch.nonzero(torch.as_tensor((t.type_flags[t.slot_type[:D]] & 1) == 1)).flatten().tolist():
        members = in_plate & (cid == p)
        n = members.sum(0)
        assert bool((torch.where(members, pos, torch.zeros_like(pos)).sum(0) == n * (n - 1) // 2).all())
        last = members & (pos == (n - 1)[None, :])
        assert bool((free[members] == last[members].to(free.dtype)).all())
    # static content: at most two items per cell (Cutboard + spawned Bread), the top one is free
    in_static = (pres == 1) & (ck == 1)
    cell = o & 63
    for c in range(64):
        m = in_static & (cell == c)
        n = m.sum(0)
        if int(n.max()) == 0:
            continue
        assert int(n.max()) <= 2
        last = m & (pos == (n - 1)[None, :])
        assert bool((last.sum(0) == (n > 0)).all())
        assert bool((free[m] == last[m].to(free.dtype)).all())
    # conservation: nothing but Bread changes its population (Bread.chop spawns a twin)
    return pres


def test_fullsize_sampled_lockstep_invariants_and_idempotence():
    env = _env()
    rid = _recipe_ids(N, 1)
    lids = env.default_layout_ids().cpu().numpy()
    obs0 = env.reset(layout_ids=lids, recipe_ids=rid).clone()
    pres0 = _check_invariants(env).clone()
    picks = np.linspace(0, N - 1, 48).astype(int)
    oracles = {int(k): OracleEnv(env.tables.layouts[lids[k]], [BOOK[int(r)] for r in rid[k]], 400,
                                 end_condition_all_dishes=True) for k in picks}
    for k, orc in oracles.items():
        assert_obs_equal(np.stack([orc.observe(i) for i in range(2)]), obs0[k].cpu().numpy(), f"env {k} reset")
    acts = _actions(60, N, 2)
    for t in range(60):
        obs, rew, term, trunc, _ = env.step(acts[t])
        a = acts[t].cpu().numpy()
        o, r = obs[picks].cpu().numpy(), rew[picks].cpu().numpy()
        for j, k in enumerate(picks):
            rr, te, tu, _ = oracles[int(k)].step(a[k])
            assert np.array_equal(bits(rr), bits(r[j])), (k, t)
            assert_obs_equal(np.stack([oracles[int(k)].observe(i) for i in range(2)]), o[j], f"env {k} step {t}")
    pres = _check_invariants(env)
    tb = env.tables
    bread = slice(int(tb.type_base[tb.dyn_types.index("Bread")]), int(tb.type_base[tb.dyn_types.index("Bread")]) + 4)
    keep = torch.ones(tb.num_dyn_slots, dtype=torch.bool, device=pres.device)
    keep[bread] = False
    assert torch.equal(pres[keep], pres0[keep])                       # populations conserved
    assert bool((pres[bread].sum(0) >= pres0[bread].sum(0)).all())     # Bread only ever multiplies
    assert int(env.error_flags.abs().sum()) == 0
    stepped = env.obs.clone()
    env.obs.fill_(float("nan"))                                       # observe() must rewrite every element
    assert torch.equal(stepped.view(torch.int64), env.observe().view(torch.int64))   # obs == f(state)


def test_fullsize_determinism_and_halves():
    """same seeds -> bit-identical outputs; the two halves stepped as separate shards reproduce the whole"""
    rid = _recipe_ids(N, 3)
    acts = _actions(25, N, 4)
    runs = []
    for _ in range(2):
        env = _env(auto_reset=True, seed=5)
        env.reset(recipe_ids=rid)
        for t in range(25):
            obs, rew, term, trunc, _ = env.step(acts[t])
        runs.append((obs.clone(), rew.clone(), env.state.clone()))
        del env
    assert torch.equal(runs[0][0].view(torch.int64), runs[1][0].view(torch.int64))
    assert torch.equal(runs[0][2], runs[1][2])
    h = N // 2
    lo = _env(h, auto_reset=True, seed=5)
    hi = _env(h, auto_reset=True, seed=5, env_offset=h)
    lo.reset(recipe_ids=rid[:h]); hi.reset(recipe_ids=rid[h:])
    for t in range(25):
        o1, r1, *_ = lo.step(acts[t, :h].contiguous())
        o2, r2, *_ = hi.step(acts[t, h:].contiguous())
    assert torch.equal(torch.cat([o1, o2]).view(torch.int64), runs[0][0].view(torch.int64))
    assert torch.equal(torch.cat([r1, r2]).view(torch.int64), runs[0][1].view(torch.int64))


def test_fullsize_exhaustive_lockstep_against_the_compiled_oracle():
    """every one of the 131072 environments, every step: observations (f64), rewards (f64) and flags of the CUDA
    path against oracle/cz_oracle.c (itself pinned to the reference's golden traces by tests/test_c_oracle.py)"""
    from oracle.cz_oracle_c import CBatch
    env = _env()
    rid = _recipe_ids(N, 21)
    lids = env.default_layout_ids().cpu().numpy()
    obs = env.reset(layout_ids=lids, recipe_ids=rid).cpu().numpy()
    ridn = rid.numpy()
    cpu = CBatch([env.tables.layouts[l] for l in lids], [[BOOK[int(r)] for r in row] for row in ridn], 400,
                 end_condition_all_dishes=True)
    assert np.array_equal(bits(cpu.observe()), bits(obs))
    acts = _actions(40, N, 22)
    for t in range(40):
        o, r, te, tu, _ = env.step(acts[t])
        co, cr, cte, ctu = cpu.step(acts[t].cpu().numpy())
        assert np.array_equal(bits(cr), bits(r.cpu().numpy())), t
        assert np.array_equal(cte, te.cpu().numpy()) and np.array_equal(ctu, tu.cpu().numpy()), t
        got = o.cpu().numpy()
        if not np.array_equal(bits(co), bits(got)):
            bad = np.argwhere(bits(co) != bits(got))
            raise AssertionError(f"step {t}: obs differ at {bad[:5].tolist()}")
    assert int(env.error_flags.abs().sum()) == 0


@pytest.mark.parametrize("name,level,meta,A,recipes,scheme,spawn", [
    ("scheme1_spawn_open4", "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", 4,
     ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"], "scheme1", (0.2, 0.15, 3)),
    ("scheme3_spawn_coop", "coop_test", "example", 2, ["TomatoLettuceSalad", "CarrotBanana"], "scheme3", (0.3, 0.1, 2)),
    ("scheme1_switch", "switch_test", "example", 2, ["TomatoLettuceSalad", "CarrotBanana"], "scheme1", None),
    ("scheme3_tiny4", "tests/golden/levels/tiny4.json", "tests/golden/levels/meta4.json", 3,
     ["TomatoSalad", "no_recipe", "no_recipe"], "scheme3", None),
])
def test_exhaustive_lockstep_other_paths(name, level, meta, A, recipes, scheme, spawn):
    """16384 environments x 80 steps, every environment compared every step: scheme1, despawn/respawn from the
    shared stream, 3/4 agents (generic and NA-specialised kernels), Switch/Block — CUDA vs oracle/cz_oracle.c"""
    import os
    from cooking_zoo_b200 import BatchedCookingEnv
    from oracle.cz_oracle import SpawnStream
    from oracle.cz_oracle_c import CBatch
    from tests.replay import ROOT
    n = 16384
    lv = os.path.join(ROOT, level) if level.endswith(".json") else level
    mt = os.path.join(ROOT, meta) if meta.endswith(".json") else meta
    kw = {} if not spawn else dict(agent_respawn_rate=spawn[0], agent_despawn_rate=spawn[1], grace_period=spawn[2])
    env = BatchedCookingEnv(n, lv, mt, A, 10 ** 5, recipes, end_condition_all_dishes=True, action_scheme=scheme,
                            layout_pool_size=64, layout_seed=9, seed=31, **kw)
    lids = env.default_layout_ids().cpu().numpy()
    obs = env.reset(layout_ids=lids).cpu().numpy()
    cpu = CBatch([env.tables.layouts[l] for l in lids], [recipes] * n, 10 ** 5, end_condition_all_dishes=True,
                 action_scheme=scheme, **kw)
    if spawn:
        for k, e in enumerate(cpu.envs):
            cpu.lib.czo_set_stream(e.h, 31, k, 1)
    assert np.array_equal(bits(cpu.observe()), bits(obs))
    g = torch.Generator(device="cpu").manual_seed(77)
    hi = 8 if scheme == "scheme1" else 5
    for t in range(80):
        act = torch.randint(0, hi, (n, A), generator=g, dtype=torch.uint8)
        o, r, te, tu, _ = env.step(act.cuda())
        co, cr, cte, ctu = cpu.step(act.numpy())
        assert np.array_equal(bits(cr), bits(r.cpu().numpy())), (name, t)
        assert np.array_equal(cte, te.cpu().numpy()) and np.array_equal(ctu, tu.cpu().numpy()), (name, t)
        assert np.array_equal(bits(co), bits(o.cpu().numpy())), (name, t)
    assert int(env.error_flags.abs().sum()) == 0


def test_config4_whole_population_on_one_gpu():
    """Maximum size: BASELINE config 4's whole population (1048576 environments) as ONE batch on one GPU — the observation
    buffer is 4.66 GB, so every row offset is 64-bit arithmetic.  Checks: structural invariants of the final state, a
    131072-environment shard stepped alone (env_offset) reproduces its slice bit for bit, sampled environments (first, last,
    around the 2^32-byte boundary of the rows) in lockstep with the oracle, observe() == the rows of the last step."""
    n, shard, steps = 1 << 20, 5, 12
    rid = _recipe_ids(n, 8)
    g = torch.Generator(device="cpu").manual_seed(21)
    acts = torch.randint(0, 5, (steps, n, 2), generator=g, dtype=torch.uint8).cuda()
    env = _env(n, auto_reset=True, seed=9)
    lids = env.default_layout_ids().cpu().numpy()
    env.reset(layout_ids=lids, recipe_ids=rid)
    lo, hi = shard * N, (shard + 1) * N
    part = _env(N, auto_reset=True, seed=9, env_offset=lo)
    assert np.array_equal(part.default_layout_ids().cpu().numpy(), lids[lo:hi])
    part.reset(layout_ids=lids[lo:hi], recipe_ids=rid[lo:hi])
    assert torch.equal(env.obs[lo:hi].view(torch.int64), part.obs.view(torch.int64))
    row_bytes = 2 * env.obs_len * 8
    edge = (1 << 32) // row_bytes                      # the environment whose rows straddle byte 2^32 of the buffer
    picks = sorted({0, 1, n - 1, n - 2, edge - 1, edge, edge + 1, lo, hi - 1, 777_777})
    oracles = {k: OracleEnv(env.tables.layouts[lids[k]], [BOOK[int(r)] for r in rid[k]], 400, end_condition_all_dishes=True)
               for k in picks}
    idx = torch.as_tensor(picks).cuda()
    for t in range(steps):
        obs, rew, term, trunc, _ = env.step(acts[t])
        o2, r2, te2, tr2, _ = part.step(acts[t, lo:hi].contiguous())
        assert torch.equal(obs[lo:hi].view(torch.int64), o2.view(torch.int64)), t
        assert torch.equal(rew[lo:hi].view(torch.int64), r2.view(torch.int64)) and torch.equal(term[lo:hi], te2), t
        o, r = obs[idx].cpu().numpy(), rew[idx].cpu().numpy()
        a = acts[t][idx].cpu().numpy()
        for j, k in enumerate(picks):
            rr, te, tu, _ = oracles[k].step(a[j])
            assert np.array_equal(bits(rr), bits(r[j])), (k, t)
            assert_obs_equal(np.stack([oracles[k].observe(i) for i in range(2)]), o[j], f"env {k} step {t}")
    assert torch.equal(env.state[:, lo:hi], part.state)
    assert int(env.error_flags.abs().sum()) == 0
    last = obs[idx].clone()
    env.obs.fill_(float("nan"))
    assert torch.equal(env.observe()[idx].view(torch.int64), last.view(torch.int64))
    del part
    torch.cuda.empty_cache()
    _check_invariants(env)


@pytest.mark.parametrize("mode", ["pipelined", "f32", "k_steps"])
def test_config4_whole_population_other_modes_equal_the_in_place_step(mode):
    """1048576 environments in one batch: the pipelined step, the float32 rows and the K-steps-per-call entry point against
    the in-place float64 step (the same 64-bit row offsets in every writer)"""
    n, steps = 1 << 20, 6
    rid = _recipe_ids(n, 8)
    g = torch.Generator(device="cpu").manual_seed(22)
    acts = torch.randint(0, 5, (steps, n, 2), generator=g, dtype=torch.uint8).cuda()
    a = _env(n, auto_reset=True, seed=9)
    kw = {"pipelined": dict(pipelined=True), "f32": dict(obs_dtype=torch.float32), "k_steps": {}}[mode]
    b = _env(n, auto_reset=True, seed=9, **kw)
    a.reset(recipe_ids=rid); b.reset(recipe_ids=rid)
    for t in range(steps):
        oa, ra, ta, ua, _ = a.step(acts[t])
        if mode != "k_steps":
            ob, rb, tb, ub, _ = b.step(acts[t])
            b.wait()
    if mode == "k_steps":
        ob, rb, tb, ub, _ = b.step_k(steps, actions=acts)
    if mode == "f32":
        for lo in range(0, n, 1 << 18):      # compare in slices: oa.float() of the whole buffer would be another 2.3 GB
            assert torch.equal(oa[lo:lo + (1 << 18)].float().view(torch.int32), ob[lo:lo + (1 << 18)].view(torch.int32)), lo
    else:
        assert torch.equal(oa.view(torch.int64), ob.view(torch.int64))
    assert torch.equal(ra.view(torch.int64), rb.view(torch.int64)) and torch.equal(ta, tb) and torch.equal(ua, ub)
    assert torch.equal(a.state, b.state)


def test_single_agent_large_batch_every_environment_against_the_compiled_oracle():
    """BASELINE config 1's shape (one agent, coop_test, TomatoLettuceSalad) as a large batch: 50001 environments (odd: the
    two-environments-per-warp row writer ends on a lone environment), every environment compared with oracle/cz_oracle.c
    every step; observe() rewrites the same rows."""
    from cooking_zoo_b200 import BatchedCookingEnv
    from oracle.cz_oracle_c import CBatch
    n, recipes = 50001, ["TomatoLettuceSalad"]
    env = BatchedCookingEnv(n, "coop_test", "example", 1, 10 ** 5, recipes, action_scheme="scheme3", layout_pool_size=64,
                            layout_seed=4, seed=12)
    lids = env.default_layout_ids().cpu().numpy()
    obs = env.reset(layout_ids=lids).cpu().numpy()
    cpu = CBatch([env.tables.layouts[l] for l in lids], [recipes] * n, 10 ** 5, action_scheme="scheme3")
    assert np.array_equal(bits(cpu.observe()), bits(obs))
    g = torch.Generator(device="cpu").manual_seed(78)
    for t in range(40):
        act = torch.randint(0, 5, (n, 1), generator=g, dtype=torch.uint8)
        o, r, te, tu, _ = env.step(act.cuda())
        co, cr, cte, ctu = cpu.step(act.numpy())
        assert np.array_equal(bits(cr), bits(r.cpu().numpy())), t
        assert np.array_equal(cte, te.cpu().numpy()) and np.array_equal(ctu, tu.cpu().numpy()), t
        assert np.array_equal(bits(co), bits(o.cpu().numpy())), t
    last = env.obs.clone()
    env.obs.fill_(float("nan"))
    assert torch.equal(env.observe().view(torch.int64), last.view(torch.int64))
    assert int(env.error_flags.abs().sum()) == 0


@pytest.mark.parametrize("name,level,A,n", [
    ("open4_a3_two_pairs_per_lane", "tests/golden/levels/open4.json", 3, 5003),
    ("open4_a4_two_pairs_per_lane", "tests/golden/levels/open4.json", 4, 5003),
    ("tiny4_a3", "tests/golden/levels/tiny4.json", 3, 24577),
    ("tiny4_a4", "tests/golden/levels/tiny4.json", 4, 24577),
    ("open4_a2", "tests/golden/levels/open4.json", 2, 24577),
])
def test_whole_row_writer_instantiations_against_the_compiled_oracle(name, level, A, n):
    """every instantiation of the whole-row TMA writer (2-4 agents, one or two (observer, slot) pairs per lane) on a batch
    large enough for the two-launch in-place step, ragged last block, every environment compared with oracle/cz_oracle.c"""
    import os
    from cooking_zoo_b200 import BatchedCookingEnv
    from oracle.cz_oracle_c import CBatch
    from tests.replay import ROOT
    recipes = ["TomatoSalad", "no_recipe", "TomatoLettuceSalad", "CarrotBanana"][:A]
    lv, mt = os.path.join(ROOT, level), os.path.join(ROOT, "tests/golden/levels/meta4.json")
    env = BatchedCookingEnv(n, lv, mt, A, 10 ** 5, recipes, end_condition_all_dishes=True, action_scheme="scheme3",
                            layout_pool_size=32, layout_seed=2, seed=41)
    lids = env.default_layout_ids().cpu().numpy()
    obs = env.reset(layout_ids=lids).cpu().numpy()
    cpu = CBatch([env.tables.layouts[l] for l in lids], [recipes] * n, 10 ** 5, end_condition_all_dishes=True, action_scheme="scheme3")
    assert np.array_equal(bits(cpu.observe()), bits(obs))
    g = torch.Generator(device="cpu").manual_seed(79)
    for t in range(30):
        act = torch.randint(0, 5, (n, A), generator=g, dtype=torch.uint8)
        l0 = env.lib.cz_launch_count()
        o, r, te, tu, _ = env.step(act.cuda())
        assert env.lib.cz_launch_count() - l0 == 2, name          # dynamics kernel + row writer
        co, cr, cte, ctu = cpu.step(act.numpy())
        assert np.array_equal(bits(cr), bits(r.cpu().numpy())), (name, t)
        assert np.array_equal(cte, te.cpu().numpy()) and np.array_equal(ctu, tu.cpu().numpy()), (name, t)
        assert np.array_equal(bits(co), bits(o.cpu().numpy())), (name, t)
    assert int(env.error_flags.abs().sum()) == 0
