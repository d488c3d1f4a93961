"""In-place step: single fused launch vs dynamics kernel + row-writer kernel, over batch sizes (CZ_TWO_KERNEL_MIN_ENVS)."""
import os, subprocess, sys
CODE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv
n = int(sys.argv[1])
env = BatchedCookingEnv(n, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                        action_scheme="scheme3", layout_pool_size=400, auto_reset=True, seed=7)
env.reset()
act = torch.randint(0, 5, (16, n, 2), dtype=torch.uint8, device="cuda")
for s in range(50): env.step(act[s % 16])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k = 1000
a.record()
for s in range(k): env.step(act[s % 16])
b.record(); torch.cuda.synchronize()
print(f"n={n} two_kernel_min={os.environ['CZ_TWO_KERNEL_MIN_ENVS']} {a.elapsed_time(b)/k*1e3:.1f} us/step {n*k/a.elapsed_time(b)/1e3:.1f} M/s")
'''
for n in (4096, 8192, 16384, 32768, 65536, 131072):
    for m in ("0", "1"):
        env = dict(os.environ, CZ_TWO_KERNEL_MIN_ENVS=m)
        print(subprocess.run([sys.executable, "-c", CODE, str(n)], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
