"""Development aid: a short Hello World run for ncu (SURVEY 8(d) actions: uniform 0..3 + 1 % quit).  python scripts/hello_prof.py N T"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
from scripts.hello_time import survey_actions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = NativeGame(expected_spec("hello", max_episode_steps=100, track_returns=True), n)
outs = g.alloc_outputs(T, discount=True)
acts = survey_actions(g, T, 1)
for i in range(6):
    g.rollout(acts, *outs)
torch.cuda.synchronize()
