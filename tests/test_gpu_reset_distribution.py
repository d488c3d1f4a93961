"""GPU: reset / auto-reset follow the reference's episode-start distribution: the default layout
pool is the exact support of the level parser with its probabilities (layout.enumerate_layouts, pinned against the
parser itself in tests/test_host_logic.py); here the DEVICE draws are tested against those probabilities."""
import os

import numpy as np
import pytest
import torch

from oracle.cz_oracle import OracleEnv
from tests.replay import ROOT, assert_obs_equal
from tests.test_host_logic import _chi2_ok

pytestmark = pytest.mark.gpu
R2 = ["TomatoLettuceSalad", "CarrotBanana"]
WEIGHTED = os.path.join(ROOT, "tests/golden/levels/weighted_layouts.json")


def _env(n, level, A, **kw):
    from cooking_zoo_b200 import BatchedCookingEnv
    return BatchedCookingEnv(n, level, "example", A, kw.pop("max_steps", 400), R2[:max(A, 1)] if A > 1 else ["TomatoSalad"],
                             end_condition_all_dishes=False, action_scheme="scheme3", **kw)


@pytest.mark.parametrize("level,A,n_layouts", [("coop_test", 2, 400), (WEIGHTED, 1, 960)])
def test_a_million_device_draws_follow_the_exact_layout_distribution(level, A, n_layouts):
    n = 1 << 18
    env = _env(n, level, A, seed=123, env_offset=5_000_000)
    t = env.tables
    assert t.layout_exact and t.num_layouts == n_layouts
    counts = np.zeros(n_layouts, np.int64)
    for episode in range(4):                       # 4 x 262144 = 1048576 draws, the ones CZ_STEP_AUTO_RESET makes
        ids = env.default_layout_ids(episode).cpu().numpy()
        counts += np.bincount(ids, minlength=n_layouts)
        for e in (0, 1, 77777, n - 1):             # host twin of the device rule
            assert env.lib.cz_layout_index(env._handle, 123, 5_000_000 + e, episode) == ids[e]
    ok, chi2, dof = _chi2_ok(counts, t.layout_prob, 4 * n)
    assert ok, (chi2, dof)
    assert counts.min() > 0                        # every layout of the support is reachable


@pytest.mark.parametrize("kernel", ["warp", "lane"])
def test_auto_reset_draws_from_the_weighted_pool(kernel, monkeypatch):
    if kernel == "lane":
        monkeypatch.setenv("CZ_WARP_MAX_ENVS", "0")
    n = 3000
    env = _env(n, WEIGHTED, 1, seed=9, auto_reset=True, max_steps=5)
    env.reset()
    assert torch.equal(env.state[-1], torch.ones(n, dtype=torch.int32, device="cuda"))          # episode counter
    rng = np.random.default_rng(1)
    for t in range(12):                            # time-up on steps 5 and 11; the step after re-initialises
        obs, rew, te, tr, _ = env.step(torch.from_numpy(rng.integers(0, 5, size=(n, 1)).astype(np.uint8)))
        if t in (5, 11):
            episode = 1 if t == 5 else 2
            want = env.default_layout_ids(episode).cpu().numpy()
            o = obs.cpu().numpy()
            for k in (0, 1, 1499, n - 1):
                orc = OracleEnv(env.tables.layouts[want[k]], ["TomatoSalad"], 5)
                assert_obs_equal(np.stack([orc.observe(0)]), o[k], f"auto-reset env {k} episode {episode}")
    assert int(env.error_flags.abs().sum()) == 0
