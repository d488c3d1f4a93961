"""GPU: float32 observation mode (reported separately).  The bar: obs32 is bit-for-bit
reference_obs.astype(float32); state, rewards and flags are untouched by the mode."""
import numpy as np
import pytest
import torch

from tests.replay import golden_files, load_golden, bits
from tests.test_gpu_parity import _make

pytestmark = pytest.mark.gpu
PICK = ("cfg1_uniform", "cfg2_sticky", "book_4", "heuristic_all", "switch_uniform", "open4_agents4", "open4_agents3",
        "tiny4_agents4", "scheme1_cfg2", "coexistence_sticky", "spawn_open4", "policy_cfg5_a1")


def f32bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


@pytest.mark.parametrize("path", [p for p in golden_files() if p.split("/")[-1][:-4] in PICK],
                         ids=lambda p: p.split("/")[-1][:-4])
def test_f32_rows_are_the_rounded_reference_rows(path):
    g = load_golden(path)
    cfg = g["config"]
    n = len(g["layouts"])
    env = _make(n, cfg, layouts=g["layouts"], obs_dtype=torch.float32)
    obs = env.reset(layout_ids=np.arange(n))
    assert obs.dtype == torch.float32
    obs = obs.cpu().numpy()
    for k in range(n):
        assert np.array_equal(f32bits(g["obs"][k, 0]), f32bits(obs[k])), f"{path} trace {k} reset"
    for t in range(g["actions"].shape[1]):
        live = [k for k in range(n) if t < g["length"][k]]
        if not live:
            break
        for k in live:
            for i in range(cfg["num_agents"]):
                if g["teleport"][k, t, i, 0] >= 0:
                    env.teleport(k, i, *map(int, g["teleport"][k, t, i]))
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(g["actions"][:, t].astype(np.uint8)))
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for k in live:
            ctx = f"{path} trace {k} step {t}"
            assert np.array_equal(f32bits(g["obs"][k, t + 1]), f32bits(obs[k])), ctx
            assert np.array_equal(bits(g["reward"][k, t]), bits(rew[k])), ctx
            assert np.array_equal(g["term"][k, t], term[k]) and np.array_equal(g["trunc"][k, t], trunc[k]), ctx
    assert np.array_equal(f32bits(env.observe().cpu().numpy()), f32bits(obs))


@pytest.mark.parametrize("pipelined", [False, True])
def test_f32_mode_tracks_the_f64_mode_at_scale(pipelined):
    """40001 auto-resetting envs, 60 steps: identical rewards / flags / state, obs32 == obs64.float() bit for bit
    (large batches: two environments per writer warp, the odd count leaves a lone last environment)"""
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=23,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    n = 40001
    a = _make(n, cfg, auto_reset=True, seed=5, layout_pool_size=64)
    b = _make(n, cfg, auto_reset=True, seed=5, layout_pool_size=64, obs_dtype=torch.float32, pipelined=pipelined)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa.float().view(torch.int32), ob.view(torch.int32))
    rng = np.random.default_rng(9)
    for t in range(60):
        act = torch.from_numpy(rng.integers(0, 5, size=(n, 2)).astype(np.uint8)).cuda()
        oa, ra, ta, ua, _ = a.step(act)
        ob, rb, tb, ub, _ = b.step(act)
        b.wait()
        assert torch.equal(oa.float().view(torch.int32), ob.view(torch.int32)), t
        assert torch.equal(ra.view(torch.int64), rb.view(torch.int64)) and torch.equal(ta, tb) and torch.equal(ua, ub), t
    assert torch.equal(a.state, b.state)


@pytest.mark.parametrize("agents,n", [(1, 5001), (2, 2368), (3, 3001)])
def test_f32_writers_by_agent_count(agents, n):
    """one agent (two environments per warp, 1112-byte rows), the smallest two-per-warp batch, three agents (one per warp)"""
    if agents == 3:
        cfg = dict(level="tests/golden/levels/open4.json", meta_file="tests/golden/levels/meta4.json", num_agents=3, max_steps=19,
                   recipes=["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad"], end_all=True, reward_scheme=None)
    else:
        cfg = dict(level="coop_test", meta_file="example", num_agents=agents, max_steps=19,
                   recipes=["TomatoLettuceSalad", "CarrotBanana"][:agents], end_all=True, reward_scheme=None)
    a = _make(n, cfg, auto_reset=True, seed=2, layout_pool_size=32)
    b = _make(n, cfg, auto_reset=True, seed=2, layout_pool_size=32, obs_dtype=torch.float32)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa.float().view(torch.int32), ob.view(torch.int32))
    rng = np.random.default_rng(4)
    for t in range(25):
        act = torch.from_numpy(rng.integers(0, 5, size=(n, agents)).astype(np.uint8)).cuda()
        oa, ob = a.step(act)[0], b.step(act)[0]
        assert torch.equal(oa.float().view(torch.int32), ob.view(torch.int32)), t
        assert torch.equal(b.observe().view(torch.int32), ob.view(torch.int32)), t
    assert torch.equal(a.state, b.state)


def test_f32_host_step_and_null_obs():
    import ctypes as C
    from cooking_zoo_b200 import _native
    cfg = dict(level="coop_test", meta_file="example", num_agents=2, max_steps=50,
               recipes=["TomatoLettuceSalad", "CarrotBanana"], end_all=True, reward_scheme=None)
    n = 333
    a = _make(n, cfg)
    b = _make(n, cfg)
    a.reset(); b.reset()
    L = a.obs_len
    act = torch.randint(0, 5, (n, 2), dtype=torch.uint8)
    h_obs = torch.empty((n, 2, L), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((n, 2), dtype=torch.float64).pin_memory()
    h_te = torch.empty((n, 2), dtype=torch.uint8).pin_memory()
    h_tr = torch.empty((n, 2), dtype=torch.uint8).pin_memory()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(b.lib.cz_step_host(b._handle, b.state.data_ptr(), act.data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(),
                                     h_te.data_ptr(), h_tr.data_ptr(), n, _native.STEP_OBS_F32, 0, 0, stream))
    oa, ra, *_ = a.step(act)
    assert torch.equal(oa.float().cpu().view(torch.int32), h_obs.view(torch.int32))
    assert torch.equal(ra.cpu(), h_rew)
    # obs == NULL: dynamics only, the observation buffer is not touched
    before = b.obs.clone()
    _native.check(b.lib.cz_step(b._handle, b.state.data_ptr(), act.cuda().data_ptr(), None, b.reward.data_ptr(),
                                b.terminated.data_ptr(), b.truncated.data_ptr(), None, n, 1, 0, 0, 0, 0, stream))
    a.step(act)
    torch.cuda.synchronize()
    assert torch.equal(b.obs, before) and torch.equal(a.state, b.state)
