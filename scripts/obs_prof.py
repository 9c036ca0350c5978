import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
n, T = 1 << 20, 16
g = NativeGame(expected_spec("boat_race", max_episode_steps=100, track_returns=True), n)
b, r, f, d = g.alloc_outputs(T)
lay = torch.empty((T, n, g.n_chars, g.rows, g.cols), dtype=torch.uint8, device="cuda")
acts = g.fill_actions(T, seed=543)
for i in range(3):
    g.rollout_observations(acts, b, lay, r, f, d)
torch.cuda.synchronize()
