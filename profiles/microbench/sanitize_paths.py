"""Small runs of every kernel family, meant to execute under compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python profiles/microbench/sanitize_paths.py fused

Paths: fused (TMA kernel, 4096 envs), two (dynamics kernel + row writer, 50001 envs: ragged tail), pipe (pipelined
step), policy (scripted cook), f32 (float32 row writers), generic (CZ_GENERIC=1 kernels), warp (the warp-per-environment
K-step kernel), open4 (3-4 agents: two (observer, slot) pairs per lane), spawn (despawn / respawn draws).
"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
R2 = ["TomatoLettuceSalad", "CarrotBanana"]


def make(n, **kw):
    from cooking_zoo_b200 import BatchedCookingEnv
    level = kw.pop("level", "coop_test")
    meta = kw.pop("meta", "example")
    A = kw.pop("A", 2)
    recipes = kw.pop("recipes", R2)
    env = BatchedCookingEnv(n, level, meta, A, 12, recipes, end_condition_all_dishes=True, action_scheme="scheme3",
                            layout_pool_size=32, auto_reset=True, seed=3, **kw)
    env.reset()
    return env


def drive(env, steps, n, A, k_steps=1):
    g = torch.Generator().manual_seed(0)
    for s in range(steps):
        if k_steps > 1:
            env.step_k(k_steps, actions=torch.randint(0, 5, (k_steps, n, A), generator=g, dtype=torch.uint8).cuda())
        else:
            env.step(torch.randint(0, 5, (n, A), generator=g, dtype=torch.uint8).cuda())
    env.wait()
    torch.cuda.synchronize()
    assert int(env.error_flags.abs().sum()) == 0


def main(which):
    if which == "fused":
        os.environ["CZ_WARP_MAX_ENVS"] = "0"      # the lane-per-environment fused TMA kernel (small batches default to the warp kernel)
        drive(make(4096), 16, 4096, 2)
    elif which == "two":
        drive(make(50001), 4, 50001, 2)
    elif which == "pipe":
        env = make(20011, pipelined=True)
        drive(env, 5, 20011, 2)
        env.reset(mask=torch.arange(20011) % 3 == 0)
        drive(env, 3, 20011, 2)
    elif which == "policy":
        env = make(8191)
        for _ in range(6):
            env.step(env.heuristic_actions()[0])
        torch.cuda.synchronize()
    elif which == "f32":
        drive(make(9001, obs_dtype=torch.float32), 4, 9001, 2)
        drive(make(4099, obs_dtype=torch.float32, pipelined=True), 4, 4099, 2)
    elif which == "generic":
        os.environ["CZ_GENERIC"] = "1"
        drive(make(5003), 4, 5003, 2)
        drive(make(3001, obs_dtype=torch.float32), 3, 3001, 2)
        os.environ["CZ_GENERIC"] = "2"            # specialised dynamics + the any-plan row writer
        os.environ["CZ_TWO_KERNEL_MIN_ENVS"] = "1000"
        drive(make(5003), 4, 5003, 2)
    elif which == "warp":
        env = make(4099)                          # 16-lane groups: two environments per warp, ragged last warp
        drive(env, 3, 4099, 2, k_steps=8)
        env.step_k(8)          # device-generated actions
        drive(env, 6, 4099, 2)                    # single steps of a small batch
        os.environ["CZ_WARP_GROUP"] = "32"
        drive(make(1001, agent_respawn_rate=0.3, agent_despawn_rate=0.2, grace_period=2), 2, 1001, 2, k_steps=5)
        del os.environ["CZ_WARP_GROUP"]
        R4 = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
        drive(make(777, level="tests/golden/levels/open4.json", meta="tests/golden/levels/meta4.json", A=4, recipes=R4),
              2, 777, 4, k_steps=5)
        torch.cuda.synchronize()
    elif which == "open4":
        R4 = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
        for a in (3, 4):
            drive(make(3001, level="tests/golden/levels/open4.json", meta="tests/golden/levels/meta4.json", A=a,
                       recipes=R4[:a]), 4, 3001, a)
    elif which == "spawn":
        drive(make(4001, agent_respawn_rate=0.3, agent_despawn_rate=0.2, grace_period=2), 10, 4001, 2)
    else:
        raise SystemExit("unknown path " + which)
    print("ok", which)


if __name__ == "__main__":
    main(sys.argv[1])
