"""Development aid: the 4,096-env actor-critic rollout (BASELINE config 5) with and without the fused policy kernel,
and the policy kernel alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from examples.actor_critic_batched import GraphedRollout, Policy
from examples.worlds import make_world

def timed(fn, warm, count):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(count): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / count

for fused in (True, False):
    game = make_world("boat_race", num_envs=4096, max_episode_steps=100, track_returns=True); game.its_showtime()
    pol = Policy(175).cuda()
    roll = GraphedRollout(game, pol, 100, fused_policy=fused).capture()
    ms = timed(roll.run, 20, 100)
    print("fused_policy=%s: %.3f ms per 100-step rollout, %.2f us per env-batch step, %.3e env-steps/s" % (fused, ms, ms * 10, 4096 * 100 / ms * 1e3))
nat = game.native
x = torch.rand((4096, 175), device="cuda")
out = torch.empty(4096, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    f = lambda: nat.policy_sample(x, pol.affine1.weight.t().contiguous(), pol.affine1.bias, pol.action_head.weight, pol.action_head.bias, seed=1, out=out)
    g = torch.cuda.CUDAGraph()
    f(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(50): f()
    print("cx_policy_sample alone (graph of 50): %.2f us per launch" % (timed(g.replay, 3, 20) * 1e3 / 50))
    b, r, fl = nat.alloc_outputs()[:3]
    planes = torch.empty((4096, 7, 5, 5), dtype=torch.float32, device="cuda")
    acts = nat.fill_actions(1, seed=2)[0]
    h = lambda: nat.step_observations(acts, b, planes, r, fl)
    g2 = torch.cuda.CUDAGraph()
    h(); torch.cuda.synchronize()
    with torch.cuda.graph(g2):
        for _ in range(50): h()
    print("cx_step_observations alone (graph of 50): %.2f us per launch" % (timed(g2.replay, 3, 20) * 1e3 / 50))
for persistent in (True,):
    game = make_world("boat_race", num_envs=4096, max_episode_steps=100, track_returns=True); game.its_showtime()
    pol = Policy(175).cuda()
    roll = GraphedRollout(game, pol, 100, persistent=True).capture()
    ms = timed(roll.run, 20, 100)
    print("persistent: %.3f ms per 100-step rollout, %.2f us per env-batch step, %.3e env-steps/s" % (ms, ms * 10, 4096 * 100 / ms * 1e3))
