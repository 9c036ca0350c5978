"""Ad-hoc kernel timing (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

def run(world, n, T, reps, **kw):
    g = NativeGame(expected_spec(world, **kw), n)
    nbuf = max(2, int(400e6 // (n * T * (g.cells + 6))) + 1)   # ring of output buffers > L2
    bufs = [g.alloc_outputs(T) for _ in range(nbuf)]
    acts = [g.fill_actions(T, seed=543, t0=i * T) for i in range(nbuf)]
    torch.cuda.synchronize()
    for i in range(3):
        b, r, f, d = bufs[i % nbuf]; g.rollout(acts[i % nbuf], b, r, f, d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        b, r, f, d = bufs[i % nbuf]; g.rollout(acts[i % nbuf], b, r, f, d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    steps = n * T
    balg = 1 + 4 + 1 + g.cells + (4 if d is not None else 0)
    print("%-10s n=%d T=%d track=%d: %.3f ms/launch  %.3e env-steps/s  %.1f GB/s alg (%d B/step)" % (
        world, n, T, g.tracks, ms, steps / ms * 1e3, steps * balg / ms / 1e6, balg), flush=True)

if __name__ == "__main__":
    n = 1 << 20
    for T in (1, 4, 32):
        run("boat_race", n, T, 20)
    run("boat_race", n, 32, 20, max_episode_steps=100, track_returns=True)
    run("demo1", n, 32, 20)
    run("hello", 65536, 8, 5)
    run("hello", 65536, 8, 5, max_episode_steps=100)
