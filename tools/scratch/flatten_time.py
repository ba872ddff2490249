"""Times the flatten kernel alone (CUDA events around 20 calls) on the train.py configuration, both sources."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from pgtg_b200 import PGTGVectorEnv
kw = bench.WORKLOADS["train-py"][0]
n = 262144
for src in ("int8",):
    env = PGTGVectorEnv(n, device="cuda:0", max_episode_steps=100, **kw)
    env.reset()
    a = torch.randint(0, 9, (n,), device="cuda:0", dtype=torch.int32)
    for _ in range(3):
        env.step(a); f = env.flat_observation()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f = env.flat_observation()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(src, "flatten alone %.4f ms" % ms, "dim", f.shape[1], "GB/s written %.0f" % (f.numel() * 4 / ms / 1e6))
    env.close()
