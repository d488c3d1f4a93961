import os, sys, time, torch, numpy as np
sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv
BOOK = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana","CucumberOnion", "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]
def run(name, N, level, meta, A, recipes, steps=300, **kw):
    env = BatchedCookingEnv(N, level, meta, A, 400, recipes, end_condition_all_dishes=True, action_scheme="scheme3",
                            layout_pool_size=200, auto_reset=True, seed=1, **kw)
    env.reset()
    acts = torch.randint(0, 5, (16, N, A), dtype=torch.uint8, device="cuda")
    for s in range(20): env.step(acts[s % 16])
    env.wait(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps): env.step(acts[s % 16])
    env.wait(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = env.tables
    b = A * t.obs_len * 8 + A * 11 + 2 * t.rows * 4
    print(f"{name}: V={t.num_variants} D={t.num_dyn_slots} ncomp={t.num_comp_slots} L={t.obs_len} simple={'?'}  {ms*1e3:.1f} us/step  {N/ms/1e3:.1f} M env-steps/s  {N*b/ms/1e6:.0f} GB/s")
    env.close()
N = 131072
if len(sys.argv) > 1 and sys.argv[1] == "open4f32":
    R = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
    run("open4 A=4 f32", 65536, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", 4, R, steps=5, obs_dtype=torch.float32)
    sys.exit(0)
run("coop2 fast", N, "coop_test", "example", 2, BOOK[1:3])
os.environ["CZ_GENERIC"] = "1"
run("coop2 generic", N, "coop_test", "example", 2, BOOK[1:3])
del os.environ["CZ_GENERIC"]
run("switch2", N, "switch_test", "example", 2, BOOK[1:3])
run("coexist2", N, "coexistence_test", "example", 2, BOOK[1:3])
R = ["TomatoLettuceSalad", "CarrotBanana", "TomatoSalad", "no_recipe"]
for a in (1, 2, 3, 4):
    run(f"open4 A={a}", 65536, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", a, R[:a])
run("open4 A=4 f32", 65536, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", 4, R, obs_dtype=torch.float32)
run("open4 A=4 pipelined", 65536, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", 4, R, pipelined=True)
run("open4 A=3 pipelined", 65536, "tests/golden/levels/open4.json", "tests/golden/levels/meta4.json", 3, R[:3], pipelined=True)
run("switch2 pipelined", N, "switch_test", "example", 2, BOOK[1:3], pipelined=True)
run("coexist2 pipelined", N, "coexistence_test", "example", 2, BOOK[1:3], pipelined=True)
