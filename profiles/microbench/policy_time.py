"""Times the device cook alone (cz_policy_act on a state driven by the cook itself), CUDA events around 50 launches.

    python profiles/microbench/policy_time.py [n_envs]
"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv  # noqa: E402

R2 = ["TomatoLettuceSalad", "CarrotBanana"]

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    env = BatchedCookingEnv(n, "coop_test", "example", 2, 400, R2, end_condition_all_dishes=True, action_scheme="scheme3",
                            layout_pool_size="auto", auto_reset=True, seed=1)
    env.reset()
    for _ in range(25):
        env.step(env.heuristic_actions()[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        env.heuristic_actions()
    e1.record()
    torch.cuda.synchronize()
    print(f"cz_policy_act, {n} envs: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us per launch",
          {k: v for k, v in os.environ.items() if k.startswith('CZ_')})
