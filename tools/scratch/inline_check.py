import os, sys, torch
sys.path.insert(0, "/root/repo")
from pgtg_b200 import PGTGVectorEnv
n, ticks = 50000, 30
g = torch.Generator(device="cuda:0"); g.manual_seed(3)
acts = [torch.randint(0, 9, (n,), device="cuda:0", dtype=torch.int32, generator=g) for _ in range(ticks)]
def run():
    env = PGTGVectorEnv(n, device="cuda:0", seed=9)
    env.reset()
    out = []
    for a in acts:
        obs, rew, term, trunc, info = env.step(a)
        out.append((env._t["obs_map"].clone(), rew.clone(), term.clone()))
    st = env.episode_stats(); env.close(); return out, st
a, sa = run()
os.environ["PGTG_INLINE_MAPGEN"] = "1"
b, sb = run()
ok = all(torch.equal(x, y) for ra, rb in zip(a, b) for x, y in zip(ra, rb))
print("inline == pipeline:", ok, sa["episodes"], sb["episodes"])
