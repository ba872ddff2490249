"""Long lock-step runs of the CUDA path against the C oracle in Philox mode, beyond the test suite's budget: the BASELINE
configurations at a few thousand envs for hundreds of ticks (about a million generated maps in the default case).
Usage (GPU box): [SOAK_SCALE=8] python tools/soak_vs_oracle.py [> gpurun_out/soak.log]. Test infrastructure: the oracle is the checker."""
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import philox_compare as pc  # noqa: E402
from native_env import NativeAdapter  # noqa: E402
from oracle.oracle import OracleVectorEnv  # noqa: E402

RUNS = [
    ("default", dict(), 8192, 300, False),
    ("default+final_observation", dict(), 4096, 200, True),
    ("obstacles+lights", dict(random_map_obstacle_probability=0.5, random_map_traffic_light_probability_weight=3), 4096, 200, True),
    ("sliding-nsd", bench.WORKLOADS["sliding-nsd-1M"][0], 4096, 150, True),
    ("config3 (traffic-64k)", bench.WORKLOADS["traffic-64k"][0], 4096, 120, True),
    ("train-py", dict(bench.WORKLOADS["train-py"][0], max_episode_steps=100), 1024, 150, True),
    ("config4 (large-1M)", bench.WORKLOADS["large-1M"][0], 256, 60, False),
]

SCALE = int(os.environ.get("SOAK_SCALE", "1"))  # multiplies the env counts
for name, kw, n, ticks, final in RUNS:
    n *= SCALE
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("cuda", num_envs=n, seed=2024, final_observation=final, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=2024, final_observation=final, threads=os.cpu_count() or 8, **kw)
    episodes = pc.compare(env, ora, ticks, action_seed=7, final_obs=final, state_every=50, check_obs_every=1)
    print(f"{name:28s} {n:6d} envs x {ticks:4d} ticks: bit-equal, {episodes} finished episodes, kernels [{env.raw.kernel_info()}], {time.time() - t0:.1f} s", flush=True)
    env.close()
    ora.close()
print("soak ok")
