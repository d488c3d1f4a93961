"""Pipelined step of 131072 two-agent coop_test environments: state ring size x dynamics blocks per SM
(cz_pipeline_config), eager and as one CUDA graph of 50 steps.  python profiles/microbench/pipe_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from cooking_zoo_b200 import BatchedCookingEnv, _native

N = 131072
acts = torch.randint(0, 5, (16, N, 2), dtype=torch.uint8, device="cuda")
ref = BatchedCookingEnv(N, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                        action_scheme="scheme3", auto_reset=True, seed=1)
ref.reset()
for s in range(20 + 300 + 50 + 200):
    ref.step(acts[s % 16])
torch.cuda.synchronize()
for buffers in ([2, 3, 4] if len(sys.argv) < 2 else [2]):
    for blocks in ((0, 1, 2, 3, 4, 6) if len(sys.argv) < 2 else (2, 3, 4)):
        env = BatchedCookingEnv(N, "coop_test", "example", 2, 400, ["TomatoLettuceSalad", "CarrotBanana"], end_condition_all_dishes=True,
                                action_scheme="scheme3", auto_reset=True, seed=1, pipelined=True, pipeline_buffers=buffers,
                                background_dynamics=blocks)
        env.reset()
        for s in range(20):
            env.step(acts[s % 16])
        env.wait()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(20, 320):
            env.step(acts[s % 16])
        env.wait()
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 300 * 1e3
        K = 48 if buffers == 3 else 50          # a multiple of the ring size: the graph ends on the buffer it started on
        K = 48 if buffers in (3, 4) else 50
        _native.check(env.lib.cz_pipeline_reset(env._handle, env.lib.cz_pipeline_current(env._handle)))
        g = torch.cuda.CUDAGraph(keep_graph=True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                for s in range(K):
                    env.step(acts[s % 16])
                env.wait()
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / (4 * K) * 1e3
        # the same graph instantiated with cudaGraphInstantiateFlagUseNodePriority: the kernel nodes keep the priority of the
        # stream they were captured from (torch instantiates without the flag: every node then runs at the launch stream's)
        prio = float("nan")
        try:
            from cuda.bindings import runtime as rt
            err, gexec = rt.cudaGraphInstantiateWithFlags(g.raw_cuda_graph(), rt.cudaGraphInstantiateFlags.cudaGraphInstantiateFlagUseNodePriority)
            assert err == rt.cudaError_t.cudaSuccess, err
            st = torch.cuda.current_stream().cuda_stream
            rt.cudaGraphLaunch(gexec, st)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(4):
                rt.cudaGraphLaunch(gexec, st)
            e1.record()
            torch.cuda.synchronize()
            prio = e0.elapsed_time(e1) / (4 * K) * 1e3
            rt.cudaGraphExecDestroy(gexec)
        except Exception as ex:
            print("node-priority instantiate failed:", repr(ex)[:200])
        b = 4630 * N
        print(f"buffers={buffers} dyn_blocks/SM={blocks}: eager {eager:.2f} us/step ({b / eager / 1e3 / 6550.1:.3f})   "
              f"graph {graph:.2f} us/step ({b / graph / 1e3 / 6550.1:.3f})   graph with node priorities {prio:.2f} us/step "
              f"({b / prio / 1e3 / 6550.1:.3f})", flush=True)
        env.close()
