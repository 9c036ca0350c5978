"""Development aid: cx_board_mapper_apply throughput (ObservationToArray RGB u8, ObservationToFeatureArray f32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from campx_b200 import rendering as R


class Obs(object):
    def __init__(self, board, characters):
        self.board, self.characters = board, characters


chars = " #<>A^v"
codes = torch.tensor([ord(c) for c in chars], dtype=torch.uint8, device="cuda")
rgb = {ch: (10 * i, 20 * i, 30 * i) for i, ch in enumerate(chars)}
for rows, cols, nb in ((5, 5, 1 << 22), (13, 36, 1 << 18)):
    boards = codes[torch.randint(0, len(chars), (nb, rows, cols), device="cuda")]
    obs = Obs(boards, chars)
    convs = [("rgb u8 chw", R.ObservationToArray(rgb, dtype=np.uint8, check=False), 3),
             ("rgb u8 hwc", R.ObservationToArray(rgb, dtype=np.uint8, permute=(1, 2, 0), check=False), 3),
             ("feat f32 x7 chw", R.ObservationToFeatureArray(chars), 28),
             ("feat f32 x7 hwc", R.ObservationToFeatureArray(chars, permute=(1, 2, 0)), 28),
             ("scalar i64", R.ObservationToArray({ch: ord(ch) for ch in chars}, check=False), 8)]
    for name, conv, out_b in convs:
        for _ in range(2):
            conv(obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            conv(obs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        by = nb * rows * cols * (1 + out_b)
        print("%dx%d %-16s boards=%d: %.3f ms  %.0f GB/s" % (rows, cols, name, nb, ms, by / ms / 1e6), flush=True)
    del boards
