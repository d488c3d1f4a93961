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
    assert lib.cz_reset(h, state.data_ptr(), None, None, None, obs.data_ptr(), n, None) == -1
    assert b"layout_ids" in lib.cz_last_error()
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr() + 8, n, None) == -1
    assert b"aligned" in lib.cz_last_error()
    assert lib.cz_step(h, state.data_ptr(), None, obs.data_ptr(), rew.data_ptr(), flg.data_ptr(), flg.data_ptr(), None,
                       n, 0, 0, 0, None) == -1
    assert lib.cz_step(h, None, act.data_ptr(), obs.data_ptr(), rew.data_ptr(), flg.data_ptr(), flg.data_ptr(), None,
                       n, 0, 0, 0, None) == -1
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr(), 0, None) == 0   # empty batch
    assert lib.cz_reset(h, state.data_ptr(), lid.data_ptr(), None, None, obs.data_ptr(), n, None) == 0
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
