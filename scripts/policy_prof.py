import os, sys
sys.path.insert(0, "/root/repo")
import torch
from examples.actor_critic_batched import Policy
from examples.worlds import make_world
game = make_world("boat_race", num_envs=4096, max_episode_steps=100, track_returns=True); game.its_showtime()
nat = game.native
pol = Policy(175).cuda()
x = torch.rand((4096, 175), device="cuda")
out = torch.empty(4096, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(6):
        nat.policy_sample(x, pol.affine1.weight.t().contiguous(), pol.affine1.bias, pol.action_head.weight, pol.action_head.bias, seed=1, out=out)
torch.cuda.synchronize()
