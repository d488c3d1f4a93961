"""In-place single step over batch sizes: fused kernel vs dynamics + row-writer kernels vs the warp-per-environment kernel,
eager launches and one CUDA graph of 20 steps.  Used to place warp_max_envs / two_kernel_min_envs (cz_tables_create).
    python profiles/microbench/inplace_path_sweep.py"""
import os
import subprocess
import sys

CODE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv
n, tag = int(sys.argv[1]), sys.argv[2]
env = BatchedCookingEnv(n, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                        action_scheme="scheme3", layout_pool_size=400, auto_reset=True, seed=7)
env.reset()
act = torch.randint(0, 5, (20, n, 2), dtype=torch.uint8, device="cuda")
for s in range(40): env.step(act[s % 20])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k = 400
a.record()
for s in range(k): env.step(act[s % 20])
b.record(); torch.cuda.synchronize()
eager = a.elapsed_time(b) / k * 1e3
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
        for s in range(20): env.step(act[s])
torch.cuda.current_stream().wait_stream(side)
for _ in range(3): g.replay()
torch.cuda.synchronize()
a.record()
for _ in range(20): g.replay()
b.record(); torch.cuda.synchronize()
graph = a.elapsed_time(b) / 400 * 1e3
print(f"n={n:7d} {tag:6s} eager {eager:7.2f} us/step   graph {graph:7.2f} us/step ({n / graph:7.1f} M/s, {4630 * n / graph / 1e3 / 6550.1:.3f})", flush=True)
'''
for n in (4096, 6144, 8192, 12288, 16384, 24576, 32768, 49152, 65536, 98304):
    for tag, e in (("fused", dict(CZ_TWO_KERNEL_MIN_ENVS="0", CZ_WARP_MAX_ENVS="0")),
                   ("two", dict(CZ_TWO_KERNEL_MIN_ENVS="1", CZ_WARP_MAX_ENVS="0")),
                   ("warp", dict(CZ_TWO_KERNEL_MIN_ENVS="0", CZ_WARP_MAX_ENVS="100000000"))):
        if tag == "warp" and n > 32768:
            continue
        out = subprocess.run([sys.executable, "-c", CODE, str(n), tag], env=dict(os.environ, **e), capture_output=True, text=True)
        print(out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
