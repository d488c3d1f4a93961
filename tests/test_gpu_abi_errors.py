"""GPU: error behaviour of the C ABI — codes + cz_last_error, no exceptions, no crashes."""
import ctypes as C

import numpy as np
import pytest
import torch

from cooking_zoo_b200 import _native
from cooking_zoo_b200.tables import compile_tables

pytestmark = pytest.mark.gpu


def _tables(**kw):
    return compile_tables("coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], **kw)


def test_create_rejects_bad_descriptors():
    lib = _native.load_library()
    t = _tables()
    desc, keep = _native.make_desc(t)
    h = C.c_void_p()
    assert lib.cz_tables_create(None, 0, C.byref(h)) == -1 and b"null" in lib.cz_last_error()
    desc.abi_version = 99
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == -1 and b"abi_version" in lib.cz_last_error()
    desc.abi_version = _native.ABI_VERSION
    desc.num_agents = 9
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == -3 and b"num_agents" in lib.cz_last_error()
    desc.num_agents = 2
    desc.width = 12
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == -3
    desc.width = 7
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == 0 and h.value
    assert lib.cz_state_rows(h) == t.rows
    assert lib.cz_tables_destroy(h) == 0
    assert lib.cz_tables_destroy(None) == 0


def test_step_and_reset_reject_null_and_misaligned_buffers():
    lib = _native.load_library()
    t = _tables()
    desc, keep = _native.make_desc(t)
    h = C.c_void_p()
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == 0
    n = 64
    state = torch.zeros((t.rows, n), dtype=torch.int32, device="cuda")
    obs = torch.zeros((n, 2, t.obs_len + 1), dtype=torch.float64, device="cuda")
    rew = torch.zeros((n, 2), dtype=torch.float64, device="cuda")
    flg = torch.zeros((n, 2), dtype=torch.uint8, device="cuda")
    act = torch.zeros((n, 2), dtype=torch.uint8, device="cuda")
    lid = torch.zeros((n,), dtype=torch.int32, device="cuda")
    assert lib.cz_reset(h, state.data_ptr(), None, None, None, obs.data_ptr(), None, n, None) == -1
    assert b"layout_ids" in lib.cz_last_error()
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr() + 8, None, n, None) == -1
    assert b"aligned" in lib.cz_last_error()
    assert lib.cz_step(h, state.data_ptr(), None, obs.data_ptr(), rew.data_ptr(), flg.data_ptr(), flg.data_ptr(), None,
                       n, 1, 0, 0, 0, 0, None) == -1
    assert lib.cz_step(h, None, act.data_ptr(), obs.data_ptr(), rew.data_ptr(), flg.data_ptr(), flg.data_ptr(), None,
                       n, 1, 0, 0, 0, 0, None) == -1
    assert lib.cz_step(h, state.data_ptr(), act.data_ptr(), obs.data_ptr(), rew.data_ptr(), flg.data_ptr(), flg.data_ptr(), None,
                       n, 0, 0, 0, 0, 0, None) == -1 and b"k_steps" in lib.cz_last_error()
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr(), None, 0, None) == 0   # empty batch
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr(), None, n, None) == 0
    torch.cuda.synchronize()
    assert lib.cz_launch_count() >= 1
    lib.cz_tables_destroy(h)


def test_python_side_validation():
    from cooking_zoo_b200 import BatchedCookingEnv
    env = BatchedCookingEnv(8, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"],
                            action_scheme="scheme3", layout_pool_size=4)
    with pytest.raises(ValueError):
        env.reset(layout_ids=np.full(8, 4, np.int32))           # pool has ids 0..3
    with pytest.raises(ValueError):
        env.reset(layout_ids=np.zeros(7, np.int32))
    env.reset()
    with pytest.raises(ValueError):
        env.step(np.zeros((8, 3), np.uint8))
    with pytest.raises(NotImplementedError):
        BatchedCookingEnv(8, "coop_test", "example", 2, 400, ["TomatoSalad"] * 2, action_scheme="scheme2")
    with pytest.raises(NotImplementedError):
        BatchedCookingEnv(8, "coop_test", "example", 2, 400, ["TomatoSalad"] * 2, obs_spaces=["symbolic", "feature_vector"])
    # actions outside the action space behave as a no-op (Discrete(5) under scheme3)
    o1, *_ = env.step(np.full((8, 2), 0, np.uint8))
    a = o1.clone()
    o2, *_ = env.step(np.full((8, 2), 9, np.uint8))
    assert torch.equal(a.view(torch.int64), o2.view(torch.int64))


def test_policy_and_f32_entry_points_reject_bad_arguments():
    from cooking_zoo_b200.policy import compile_policy_tables
    lib = _native.load_library()
    t = _tables()
    desc, keep = _native.make_desc(t)
    h = C.c_void_p()
    assert lib.cz_tables_create(C.byref(desc), 0, C.byref(h)) == 0
    ptab = compile_policy_tables(t)
    pdesc, pkeep = _native.make_policy_desc(t, ptab)
    p = C.c_void_p()
    assert lib.cz_policy_create(None, C.byref(pdesc), C.byref(p)) == -1
    pdesc.num_variants += 1
    assert lib.cz_policy_create(h, C.byref(pdesc), C.byref(p)) == -1 and b"num_variants" in lib.cz_last_error()
    pdesc.num_variants -= 1
    bad = ptab["first_step"].copy()
    bad[0, 9, 10] = 7                                  # not a movement action
    bdesc, bkeep = _native.make_policy_desc(t, dict(ptab, first_step=bad))
    assert lib.cz_policy_create(h, C.byref(bdesc), C.byref(p)) == -1 and b"first_step" in lib.cz_last_error()
    assert lib.cz_policy_create(h, C.byref(pdesc), C.byref(p)) == 0 and p.value
    n = 40
    state = torch.zeros((t.rows, n), dtype=torch.int32, device="cuda")
    act = torch.zeros((n, 2), dtype=torch.uint8, device="cuda")
    assert lib.cz_policy_act(p, None, None, act.data_ptr(), None, n, None) == -1
    assert lib.cz_policy_act(p, state.data_ptr(), None, None, None, n, None) == -1
    assert lib.cz_policy_act(p, state.data_ptr(), None, act.data_ptr(), None, -1, None) == -1
    assert lib.cz_policy_act(p, state.data_ptr(), None, act.data_ptr(), None, 0, None) == 0
    assert lib.cz_policy_destroy(p) == 0 and lib.cz_policy_destroy(None) == 0
    obs32 = torch.zeros((n, 2, t.obs_len + 4), dtype=torch.float32, device="cuda")
    assert lib.cz_observe_f32(h, state.data_ptr(), None, n, None) == -1
    assert lib.cz_observe_f32(h, state.data_ptr(), obs32.data_ptr() + 4, n, None) == -1 and b"aligned" in lib.cz_last_error()
    assert lib.cz_observe_f32(h, state.data_ptr(), obs32.data_ptr(), 0, None) == 0
    assert lib.cz_observe(h, state.data_ptr(), None, n, None) == -1       # the f64 observer has no NULL mode
    assert lib.cz_pipeline_wait_state(None, None) == -1
    assert lib.cz_pipeline_wait_state(h, None) == 0                      # pipeline never started: nothing to wait for
    assert lib.cz_tables_destroy(h) == 0
