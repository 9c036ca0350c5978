"""Development aid: Hello World (generic path) timing, SURVEY 8(d) workload: actions uniform over 0..3 plus 1% quit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec

def survey_actions(g, T, seed):
    a = g.fill_actions(T, seed=seed)                       # uniform over 0..4
    keep_quit = torch.rand(a.shape, device=a.device) < 0.05   # 0.2 * 0.05 = 1% quit
    repl = torch.randint(0, 4, a.shape, device=a.device, dtype=torch.uint8)
    return torch.where((a == 4) & ~keep_quit, repl, a).contiguous()

def run(n, T, reps, **kw):
    g = NativeGame(expected_spec("hello", **kw), n)
    nbuf = max(2, int(400e6 // (n * T * (g.cells + 6))) + 1)
    bufs = [g.alloc_outputs(T) for _ in range(nbuf)]
    acts = [survey_actions(g, T, 543 + i) for i in range(nbuf)]
    for i in range(3):
        b, r, f, d = bufs[i % nbuf]; g.rollout(acts[i % nbuf], b, r, f, d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        b, r, f, d = bufs[i % nbuf]; g.rollout(acts[i % nbuf], b, r, f, d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    balg = 1 + 4 + 1 + 4 + g.cells                         # action, reward, flags, discount, board
    print("hello n=%d T=%d: %.3f ms/launch  %.3e env-steps/s  %.1f GB/s alg (%d B/step)" % (
        n, T, ms, n * T / ms * 1e3, n * T * balg / ms / 1e6, balg), flush=True)

if __name__ == "__main__":
    run(65536, 8, 10, max_episode_steps=100)
    run(65536, 32, 10, max_episode_steps=100)
    run(1 << 18, 32, 5, max_episode_steps=100)
