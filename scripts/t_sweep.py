"""Development aid: per-launch fixed cost of the agent rollout kernel (time vs fused steps T)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

n = 1 << 20
g = NativeGame(expected_spec("boat_race", max_episode_steps=100, track_returns=True), n)
for T in (32, 48, 64, 96, 128):
    bufs = [g.alloc_outputs(T) for _ in range(2)]
    acts = [g.fill_actions(T, seed=543, t0=i * T) for i in range(2)]
    reps = max(4, 1024 // T)
    for i in range(3):
        b, r, f, d = bufs[i % 2]; g.rollout(acts[i % 2], b, r, f, d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        b, r, f, d = bufs[i % 2]; g.rollout(acts[i % 2], b, r, f, d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("T=%3d  %.4f ms/launch  %.4f us/step  %.1f GB/s alg" % (T, ms, ms / T * 1e3, n * (T * 31 + 14) / ms / 1e6), flush=True)
    del bufs, acts
    torch.cuda.empty_cache()
