"""Development aid: what does a pure HBM write stream reach on this GPU (vs the copy figure)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: a.fill_(7)); print("fill_ 1GiB   : %.3f ms  %.0f GB/s written" % (ms, n / ms / 1e6))
ms = timeit(lambda: a.zero_()); print("memset 1GiB  : %.3f ms  %.0f GB/s written" % (ms, n / ms / 1e6))
ms = timeit(lambda: b.copy_(a)); print("copy 1GiB    : %.3f ms  %.0f GB/s read+write" % (ms, 2 * n / ms / 1e6))
af = a.view(torch.float32)
ms = timeit(lambda: af.sum()); print("read-sum 1GiB: %.3f ms  %.0f GB/s read" % (ms, n / ms / 1e6))
from scripts.quick_time import run
for T in (32, 64, 128):
    run("boat_race", 1 << 20, T, 10)
run("boat_race", 1 << 20, 32, 10, max_episode_steps=100, track_returns=True)
run("boat_race", 1 << 21, 32, 10)
run("boat_race", 1 << 19, 32, 10)
