"""Times the row writers alone (cz_observe / cz_observe_f32 on a stepped state), CUDA events around 50 launches.

    python profiles/microbench/obs_time.py [n_envs]
"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv  # noqa: E402

R2 = ["TomatoLettuceSalad", "CarrotBanana"]


def run(n, dtype):
    env = BatchedCookingEnv(n, "coop_test", "example", 2, 400, R2, end_condition_all_dishes=True, action_scheme="scheme3",
                            layout_pool_size="auto", auto_reset=True, seed=1, obs_dtype=dtype)
    env.reset()
    g = torch.Generator().manual_seed(0)
    for _ in range(8):
        env.step(torch.randint(0, 5, (n, 2), generator=g, dtype=torch.uint8).cuda())
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
    nbytes = n * 2 * 278 * (4 if dtype == torch.float32 else 8)
    print(f"{'f32' if dtype == torch.float32 else 'f64'} rows, {n} envs: {us:.2f} us per launch, {nbytes / us / 1e3:.0f} GB/s of rows",
          {k: v for k, v in os.environ.items() if k.startswith('CZ_')})


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    for dt in (torch.float32, torch.float64):
        run(n, dt)
