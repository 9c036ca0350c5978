"""Development aid: the bench.py kernel (k_agent_rollout<TRACK>, boat_race, 2^20 envs x 32 steps) for an ncu capture."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
n, T = 1 << 20, 32
g = NativeGame(expected_spec("boat_race", max_episode_steps=100, track_returns=True), n)
bufs = [g.alloc_outputs(T) for _ in range(2)]
acts = g.fill_actions(T, seed=543)
for i in range(3):
    b, r, f, d = bufs[i % 2]
    g.rollout(acts, b, r, f, d)
torch.cuda.synchronize()
