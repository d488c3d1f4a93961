"""A few K-step launches of the warp-per-environment kernel (for ncu): python profiles/microbench/warp_loop.py [n] [K]"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["CZ_WARP_K_MAX_ENVS"] = "10000000"
env = BatchedCookingEnv(n, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                        action_scheme="scheme3", layout_pool_size=400, auto_reset=True, seed=7)
env.reset()
acts = torch.randint(0, 5, (K, n, 2), dtype=torch.uint8, device="cuda")
for _ in range(6):
    env.step_k(K, actions=acts)
torch.cuda.synchronize()
print("ok")
