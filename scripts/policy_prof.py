"""Development aid: short runs of the policy kernels for ncu.  python scripts/policy_prof.py [sample|rollout]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from examples.actor_critic_batched import Policy
from examples.worlds import make_world
game = make_world("boat_race", num_envs=4096, max_episode_steps=100, track_returns=True); game.its_showtime()
nat = game.native
pol = Policy(175).cuda()
w1t = pol.affine1.weight.detach().t().contiguous()
b1, w2, b2 = pol.affine1.bias.detach(), pol.action_head.weight.detach(), pol.action_head.bias.detach()
if (sys.argv[1:] or ["sample"])[0] == "sample":
    x = torch.rand((4096, 175), device="cuda")
    out = torch.empty(4096, dtype=torch.uint8, device="cuda")
    for _ in range(6):
        nat.policy_sample(x, w1t, b1, w2, b2, seed=1, out=out)
else:
    T = 100
    states = torch.empty((T + 1, 4096, 175), dtype=torch.float32, device="cuda")
    actions = torch.empty((T, 4096), dtype=torch.uint8, device="cuda")
    rewards = torch.empty((T, 4096), dtype=torch.float32, device="cuda")
    flags = torch.empty((T, 4096), dtype=torch.uint8, device="cuda")
    for _ in range(4):
        nat.rollout_policy(T, w1t, b1, w2, b2, 1, states, actions, rewards, flags)
torch.cuda.synchronize()
