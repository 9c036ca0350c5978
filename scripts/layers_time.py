"""Development aid: cx_layers_from_board throughput (u8 and f32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from campx_b200.runtime import NativeGame
from tests.expected_specs import expected_spec
for world, nb in (("boat_race", 1 << 24), ("hello", 1 << 20)):
    g = NativeGame(expected_spec(world), 64)
    boards = torch.randint(32, 96, (nb, g.rows, g.cols), dtype=torch.uint8, device="cuda")
    for dt in (torch.uint8, torch.float32):
        n = nb if dt == torch.uint8 else nb // 4
        out = torch.empty((n, g.n_chars, g.rows, g.cols), dtype=dt, device="cuda")
        for _ in range(2):
            g.layers_from_board(boards[:n], out=out, dtype=dt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.layers_from_board(boards[:n], out=out, dtype=dt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        by = n * g.cells * (1 + g.n_chars * out.element_size())
        print("%-10s %s boards=%d: %.3f ms  %.0f GB/s" % (world, str(dt).split(".")[1], n, ms, by / ms / 1e6), flush=True)
    del boards, out
