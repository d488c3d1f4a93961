import os, sys, torch
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "65536"]
import importlib.util
spec = importlib.util.spec_from_file_location("obs_time", "profiles/microbench/obs_time.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
for n in (65536, 131072):
    for _ in range(3):
        m.run(n, torch.float64, 1, False)
    for _ in range(2):
        m.run(n, torch.float64, 1, True)
