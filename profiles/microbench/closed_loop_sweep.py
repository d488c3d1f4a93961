"""Closed loop with the device cook, pipelined, 131072 envs: blocks per SM of the dynamics kernel x of the policy kernel.
One CUDA graph of 40 steps per setting.  python profiles/microbench/closed_loop_sweep.py
CZ_POLICY_ON_DYN=0 puts the cook's kernel back on the caller's stream (the round-2 sweep before the cook moved to the
library's high-priority stream)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv, _native

N = 131072
BOOK = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana", "CucumberOnion", "AppleWatermelon",
        "TomatoLettuceOnionSalad", "no_recipe"]
rid = torch.randint(0, 8, (N, 2), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
COMBOS = ((0, 0), (3, 0), (0, 4), (4, 4), (6, 6), (4, 8), (8, 8), (6, 10), (3, 6), (3, 3), (2, 3), (4, 5), (5, 5), (3, 4), (2, 2))
for dyn, pol in COMBOS:
    env = BatchedCookingEnv(N, "coop_test", "example", 2, 400, BOOK[1:3], end_condition_all_dishes=True, action_scheme="scheme3",
                            recipe_pool=BOOK, auto_reset=True, seed=1, pipelined=True, background_dynamics=dyn,
                            background_policy=pol)
    env.reset(recipe_ids=rid)
    for _ in range(10):
        env.step(env.heuristic_actions()[0])
    env.wait()
    torch.cuda.synchronize()
    _native.check(env.lib.cz_pipeline_reset(env._handle, env.lib.cz_pipeline_current(env._handle)))
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            for _ in range(40):
                env.step(env.heuristic_actions()[0])
            env.wait()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 160 * 1e3
    print(f"dyn blocks/SM={dyn} policy blocks/SM={pol}: {us:.2f} us/step  {N / us:.1f} M env-steps/s  ({4630 * N / us / 1e3 / 6550.1:.3f})", flush=True)
    env.close()
