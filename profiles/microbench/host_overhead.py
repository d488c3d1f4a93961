"""Host cost of one BatchedCookingEnv.step() call: 32 environments (the GPU work is negligible), 3000 calls.
    python profiles/microbench/host_overhead.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv

for pipelined in (False, True):
    env = BatchedCookingEnv(32, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                            action_scheme="scheme3", auto_reset=True, seed=1, pipelined=pipelined)
    env.reset()
    act = torch.randint(0, 5, (32, 2), dtype=torch.uint8, device="cuda")
    for _ in range(200):
        env.step(act)
    env.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3000):
        env.step(act)
    env.wait()
    torch.cuda.synchronize()
    print(f"pipelined={pipelined}: {(time.perf_counter() - t0) / 3000 * 1e6:.1f} us per step() call", flush=True)
