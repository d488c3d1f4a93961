"""In-place cz_step of 131072 two-agent coop_test environments as 1 / 2 / 3 / 4 / 8 column ranges (CZ_SPLIT), eager
launches and one CUDA graph of 50 steps: python profiles/microbench/split_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv

N = 131072
acts = torch.randint(0, 5, (16, N, 2), dtype=torch.uint8, device="cuda")
for split in ("1", "2", "3", "4", "8"):
    os.environ["CZ_SPLIT"] = split
    env = BatchedCookingEnv(N, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                            action_scheme="scheme3", auto_reset=True, seed=1)
    env.reset()
    for s in range(20):
        env.step(acts[s % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(300):
        env.step(acts[s % 16])
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 300 * 1e3
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for s in range(50):
                env.step(acts[s % 16])
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 200 * 1e3
    b = 4630 * N
    print(f"CZ_SPLIT={split}: eager {eager:.2f} us/step ({b / eager / 1e3 / 6550.1:.3f} of roofline)   graph {graph:.2f} us/step "
          f"({b / graph / 1e3 / 6550.1:.3f})")
    env.close()
