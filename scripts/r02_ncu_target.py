"""Development aid: a short run of one kernel configuration for ncu.  python scripts/r02_ncu_target.py WORLD N T [LAUNCHES]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

world, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 12
mode = sys.argv[5] if len(sys.argv) > 5 else "rollout"      # rollout | planes_f32 (cx_step_observations, T ignored)
kw = dict(max_episode_steps=100, track_returns=True)
g = NativeGame(expected_spec(world, **kw), n)
nb = 4
bufs = [g.alloc_outputs(T) for _ in range(nb)]
acts = [g.fill_actions(T, seed=1, t0=i * T) for i in range(nb)]
planes = torch.empty((n, g.n_chars, g.rows, g.cols), dtype=torch.float32, device="cuda") if mode == "planes_f32" else None
for i in range(launches):
    if planes is not None:
        b = bufs[i % nb]
        g.step_observations(acts[i % nb][0], b[0][0] if b[0].dim() == 4 else b[0], planes,
                            b[1][0] if b[1].dim() == 2 else b[1], b[2][0] if b[2].dim() == 2 else b[2])
    elif T == 1:
        g.step(acts[i % nb][0], *bufs[i % nb][:3])
    else:
        g.rollout(acts[i % nb], *bufs[i % nb])
torch.cuda.synchronize()
print("done", world, n, T)
