"""Development aid: a short run of one kernel configuration for ncu.  python scripts/r02_ncu_target.py WORLD N T [LAUNCHES]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

world, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 12
kw = dict(max_episode_steps=100, track_returns=True)
g = NativeGame(expected_spec(world, **kw), n)
nb = 4
bufs = [g.alloc_outputs(T) for _ in range(nb)]
acts = [g.fill_actions(T, seed=1, t0=i * T) for i in range(nb)]
for i in range(launches):
    if T == 1:
        g.step(acts[i % nb][0], *bufs[i % nb][:3])
    else:
        g.rollout(acts[i % nb], *bufs[i % nb])
torch.cuda.synchronize()
print("done", world, n, T)
