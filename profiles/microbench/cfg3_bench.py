"""BASELINE config 3 (4096 two-agent coop_test environments, random actions): per-launch and K-steps-per-launch rates
of the warp-per-environment kernel next to the lane-per-environment fused kernel.  python profiles/microbench/cfg3_bench.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv

R2 = ["TomatoLettuceSalad", "CarrotBanana"]


def make(n, **env):
    for k, v in env.items():
        os.environ[k] = v
    e = BatchedCookingEnv(n, "coop_test", "example", 2, 400, R2, end_condition_all_dishes=True, action_scheme="scheme3",
                          layout_pool_size=400, auto_reset=True, seed=7)
    for k in env:
        del os.environ[k]
    e.reset()
    return e


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def main():
    ns = [int(x) for x in sys.argv[1:]] or [4096]
    for n in ns:
        acts = torch.randint(0, 5, (64, n, 2), dtype=torch.uint8, device="cuda")
        lane = make(n, CZ_WARP_MAX_ENVS="0")
        warp = make(n, CZ_WARP_MAX_ENVS="1000000", CZ_WARP_K_MAX_ENVS="1000000")
        i = [0]

        def one(e):
            def f():
                e.step(acts[i[0] % 64])
                i[0] += 1
            return f
        us_lane = timed(one(lane), 1000)
        us_warp = timed(one(warp), 1000)
        print(f"n={n} per-launch: lane {us_lane:.2f} us ({n / us_lane:.1f} M/s)   warp {us_warp:.2f} us ({n / us_warp:.1f} M/s)")
        for K in (16, 64, 256):
            us_r = timed(lambda: warp.step_k(min(K, 64), actions=acts[:min(K, 64)]), 50) / min(K, 64)
            us_d = timed(lambda: warp.step_k(K), 50) / K
            print(f"n={n} K={K}: resident actions {us_r:.2f} us/step ({n / us_r:.1f} M/s)   device actions {us_d:.2f} us/step ({n / us_d:.1f} M/s)")
        lane.close()
        warp.close()


main()
