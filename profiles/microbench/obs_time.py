"""Times the row writers alone (cz_observe / cz_observe_f32 on a stepped state), CUDA events around 50 launches.

    python profiles/microbench/obs_time.py [n_envs]
"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv  # noqa: E402

R2 = ["TomatoLettuceSalad", "CarrotBanana"]


R4 = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]


def run(n, dtype, agents=2, open4=False):
    if open4:   # the 1-4 agent kitchen of config 5
        env = BatchedCookingEnv(n, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", agents, 400, R4[:agents],
                                end_condition_all_dishes=True, action_scheme="scheme3", layout_pool_size=64, auto_reset=True,
                                seed=1, obs_dtype=dtype)
    else:
        env = BatchedCookingEnv(n, "coop_test", "example", agents, 400, R2[:agents], end_condition_all_dishes=True,
                                action_scheme="scheme3", layout_pool_size="auto", auto_reset=True, seed=1, obs_dtype=dtype)
    env.reset()
    g = torch.Generator().manual_seed(0)
    for _ in range(8):
        env.step(torch.randint(0, 5, (n, agents), generator=g, dtype=torch.uint8).cuda())
    for _ in range(5):
        env.observe()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        env.observe()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    nbytes = env.obs.numel() * env.obs.element_size()
    print(f"{'f32' if dtype == torch.float32 else 'f64'} rows, {'open4' if open4 else 'coop_test'}, {agents} agents, L = {env.obs_len}, "
          f"{n} envs: {us:.2f} us per launch, {nbytes / us / 1e3:.0f} GB/s of rows",
          {k: v for k, v in os.environ.items() if k.startswith('CZ_')})


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    for dt in (torch.float32, torch.float64):
        run(n, dt)
    if os.environ.get("OBS_TIME_OPEN4"):
        for a in (1, 2, 3, 4):
            run(n, torch.float64, a, True)
        run(n, torch.float64, 1)
