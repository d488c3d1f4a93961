"""Build container only: oracle/cz_oracle.py in lockstep with the live, unmodified reference."""
import numpy as np
import pytest

from oracle.cz_oracle import OracleEnv, RECIPES
from tests.replay import assert_state_equal, assert_obs_equal, bits

pytestmark = pytest.mark.reference
BOOK = list(RECIPES)


def _lockstep(seed, level, A, recipes, max_steps, end_all, steps, policy_seed):
    from oracle.ref_harness import RefEnv
    ref = RefEnv(seed, level, "example", A, max_steps, recipes, end_condition_all_dishes=end_all)
    orc = OracleEnv(ref.layout(), recipes, max_steps, end_condition_all_dishes=end_all)
    assert_state_equal(ref.export_state(), orc.export_state(), f"seed {seed} reset")
    rng = np.random.default_rng(policy_seed)
    prev = np.zeros(A, np.int64)
    for t in range(steps):
        ctx = f"seed {seed} step {t}"
        act = np.where(rng.random(A) < 0.4, prev, rng.integers(0, 5, size=A))
        prev = act
        r1 = ref.step(act)
        r2 = orc.step(act)
        assert np.array_equal(bits(r1[0]), bits(r2[0])), ctx
        for a, b in zip(r1[1:], r2[1:]):
            assert list(a) == [int(v) for v in b], ctx
        assert_state_equal(ref.export_state(), orc.export_state(), ctx)
        assert_obs_equal(ref.observe_all(), np.stack([orc.observe(i) for i in range(A)]), ctx)
        if r1[1].any() or r1[2].any():
            break
    assert orc.error == 0


@pytest.mark.parametrize("seed", range(8))
def test_live_lockstep_coop_test(seed):
    A = 1 + seed % 2
    recipes = [BOOK[(seed + k) % len(BOOK)] for k in range(A)]
    _lockstep(seed, "coop_test", A, recipes, 150, bool(seed & 2), 150, 77 + seed)
