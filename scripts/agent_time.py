import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scripts.quick_time import run
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
run("boat_race", 1 << 20, 32, 20)
run("boat_race", 1 << 20, 32, 20, max_episode_steps=100, track_returns=True)
def run_synth(n, T, reps, write_actions, **kw):
    g = NativeGame(expected_spec("boat_race", **kw), n)
    bufs = [g.alloc_outputs(T) for _ in range(2)]
    aout = torch.empty((T, n), dtype=torch.uint8, device="cuda") if write_actions else None
    for i in range(3):
        b, r, f, d = bufs[i % 2]; g.rollout_synth(T, 543, b, r, f, d, t0=i * T, actions_out=aout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        b, r, f, d = bufs[i % 2]; g.rollout_synth(T, 543, b, r, f, d, t0=i * T, actions_out=aout)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    balg = 30 + (1 if write_actions else 0)
    print("synth track=%d actions_out=%d: %.3f ms/launch %.3e env-steps/s %.1f GB/s alg (%d B/step)" % (
        g.tracks, write_actions, ms, n * T / ms * 1e3, n * T * balg / ms / 1e6, balg), flush=True)
run_synth(1 << 20, 32, 20, False)
run_synth(1 << 20, 32, 20, True)
run_synth(1 << 20, 32, 20, True, max_episode_steps=100, track_returns=True)
