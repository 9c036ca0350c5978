"""Development aid: cx_rollout_policy (the whole actor-critic rollout in one launch) by batch size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from examples.actor_critic_batched import Policy
from examples.worlds import make_world
T = 100
for n in (4096, 16384, 65536, 262144):
    game = make_world("boat_race", num_envs=n, max_episode_steps=100, track_returns=True, verify=False); game.its_showtime()
    nat = game.native
    pol = Policy(175).cuda()
    w1t = pol.affine1.weight.detach().t().contiguous()
    b1, w2, b2 = pol.affine1.bias.detach(), pol.action_head.weight.detach(), pol.action_head.bias.detach()
    nb = 2 if n * T * 700 * 2 < 40e9 else 1
    bufs = [(torch.empty((T + 1, n, 175), dtype=torch.float32, device="cuda"), torch.empty((T, n), dtype=torch.uint8, device="cuda"),
             torch.empty((T, n), dtype=torch.float32, device="cuda"), torch.empty((T, n), dtype=torch.uint8, device="cuda")) for _ in range(nb)]
    def run(i):
        s, a, r, f = bufs[i % nb]
        nat.rollout_policy(T, w1t, b1, w2, b2, 1, s, a, r, f, step_offset=i * T)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, int(2e9 // (n * T * 731)))
    e0.record()
    for i in range(reps): run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("cx_rollout_policy n=%d T=%d: %.3f ms per rollout, %.2f us per env-batch step, %.3e env-steps/s, %.0f GB/s of states+reward+flags+actions" % (
        n, T, ms, ms * 1e3 / T, n * T / ms * 1e3, n * T * 706 / ms / 1e6), flush=True)
    del game, bufs
    torch.cuda.empty_cache()
