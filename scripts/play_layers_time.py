"""Development aid: Engine.play() + obs.layered_board every step (the full reference Observation), 2^20 envs:
fused observation step (one kernel) against step kernel + cx_layers_from_board."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from examples.worlds import make_world

def run(world, n, fused, steps=40):
    game = make_world(world, num_envs=n, max_episode_steps=100, verify=False)
    game.its_showtime()
    game.fused_observation_steps = fused
    acts = game.native.fill_actions(steps + 5, seed=3)
    obs = None
    for t in range(5):
        obs, _, _ = game.play(acts[t]); obs.layered_board
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(5, 5 + steps):
        obs, _, _ = game.play(acts[t]); obs.layered_board
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    per = 1 + 4 + 1 + game.native.cells * (1 + game.native.n_chars)
    print("%-10s n=%d fused=%-5s %.3f ms/step  %.3e env-steps/s  %.0f GB/s alg" % (
        world, n, fused, ms, n / ms * 1e3, n * per / ms / 1e6), flush=True)

if __name__ == "__main__":
    for world, n in (("boat_race", 1 << 20), ("hello", 1 << 16)):
        for fused in (False, True):
            run(world, n, fused)
