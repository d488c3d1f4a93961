import os, sys, torch
sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv
BOOK = ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana","CucumberOnion", "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]
N=131072
env = BatchedCookingEnv(N, "coop_test", "example", 2, 400, BOOK[1:3], end_condition_all_dishes=True, action_scheme="scheme3", recipe_pool=BOOK, layout_pool_size=400, auto_reset=True, seed=2026)
rid = torch.randint(0, 8, (N, 2), dtype=torch.uint8)
env.reset(recipe_ids=rid)
for _ in range(60):
    env.step(env.heuristic_actions()[0])
torch.cuda.synchronize()
