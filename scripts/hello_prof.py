import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
from scripts.hello_time import survey_actions
g = NativeGame(expected_spec("hello", max_episode_steps=100), 65536)
outs = g.alloc_outputs(16)
acts = survey_actions(g, 16, 1)
for i in range(3):
    g.rollout(acts, *outs[:3], outs[3])
torch.cuda.synchronize()
