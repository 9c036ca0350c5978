"""Development aid: single-agent worlds with boards of 97..254 cells (lane-per-env kernel), board-only contract."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from campx_b200 import things
from campx_b200.ascii_art import ascii_art_to_game, Partial
from campx_b200.runtime import NativeGame
from examples.worlds import Walker
from tests.test_gpu_generic_worlds import _random_art

def run(rows, cols, n, T, reps=8):
    art = _random_art(np.random.Generator(np.random.PCG64(rows * cols)), rows, cols, 0.2, 0.2)
    game = ascii_art_to_game(art, ' ', drapes={'A': Partial(Walker, walls='#', treasures='*'),
                                               '#': things.FixedDrape, '*': things.FixedDrape},
                             z_order='*A#', num_envs=n, max_episode_steps=100, track_returns=True)
    spec = game.compile()
    g = NativeGame(spec, n)
    bufs = [g.alloc_outputs(T) for _ in range(2)]
    acts = [g.fill_actions(T, seed=543, t0=i * T) for i in range(2)]
    for i in range(3):
        b, r, f, d = bufs[i % 2]; g.rollout(acts[i % 2], b, r, f, d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        b, r, f, d = bufs[i % 2]; g.rollout(acts[i % 2], b, r, f, d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    balg = 1 + 4 + 1 + g.cells
    print("%dx%d (%d cells, path %d) n=%d T=%d: %.3f ms/launch  %.3e env-steps/s  %.0f GB/s alg (%d B/step)" % (
        rows, cols, g.cells, g.info.path, n, T, ms, n * T / ms * 1e3, n * T * balg / ms / 1e6, balg), flush=True)

if __name__ == "__main__":
    run(8, 12, 1 << 20, 16)       # 96 cells: k_agent_rollout
    run(10, 12, 1 << 20, 16)      # 120 cells: lane-per-env kernel
    run(15, 16, 1 << 19, 16)      # 240 cells
    run(13, 21, 1 << 18, 16)      # 273 cells: generic kernels
