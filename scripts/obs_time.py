"""Development aid: timing of the fused observation kernel (board + layered board per env-step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

def run(world, n, T, reps, **kw):
    g = NativeGame(expected_spec(world, **kw), n)
    bufs = [g.alloc_outputs(T) for _ in range(2)]
    lays = [torch.empty((T, n, g.n_chars, g.rows, g.cols), dtype=torch.uint8, device="cuda") for _ in range(2)]
    acts = [g.fill_actions(T, seed=543, t0=i * T) for i in range(2)]
    def launch(i):
        b, r, f, d = bufs[i % 2]
        g.rollout_observations(acts[i % 2], b, lays[i % 2], r, f, d)
    for i in range(3):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    balg = 1 + 4 + 1 + g.cells * (1 + g.n_chars)
    print("%-10s n=%d T=%d track=%d: %.3f ms/launch  %.3e env-steps/s  %.1f GB/s alg (%d B/step)" % (
        world, n, T, g.tracks, ms, n * T / ms * 1e3, n * T * balg / ms / 1e6, balg), flush=True)
    # the two-kernel route for comparison
    def launch2(i):
        b, r, f, d = bufs[i % 2]
        g.rollout(acts[i % 2], b, r, f, d)
        g.layers_from_board(b, out=lays[i % 2])
    for i in range(2):
        launch2(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        launch2(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-10s   rollout + layers_from_board: %.3f ms  %.3e env-steps/s" % (world, ms, n * T / ms * 1e3), flush=True)

if __name__ == "__main__":
    run("boat_race", 1 << 20, 16, 10)
    run("boat_race", 1 << 20, 16, 10, max_episode_steps=100, track_returns=True)
    run("boat_race", 4096, 100, 20, max_episode_steps=100, track_returns=True)
