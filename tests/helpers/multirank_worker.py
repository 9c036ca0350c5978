"""Worker of tests/test_gpu_multirank.py, launched by torchrun with one rank per GPU.

SURVEY 8(e): "verify that an N-GPU run reproduces N single-GPU runs env-for-env".  Rank r owns envs
[r*n, (r+1)*n) of a global batch (Philox action streams keyed by the GLOBAL env id, campx_b200/dist.py); after a
rollout every rank's boards / rewards / flags are all-gathered over NCCL and rank 0 compares them, env for env, with
ONE single-GPU run of the whole world*n batch on its own device; the episode statistics go through the design's only
collective (all_reduce_stats) and must equal the single run's.  A sample of every rank's envs is also replayed through
the CPU oracle.  Prints MULTIRANK_OK on success.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from campx_b200 import _native as N
from campx_b200 import dist as cxdist
from examples.worlds import make_world
from oracle import campx_oracle as O


def main():
    n_arg, T, limit, seed = int(sys.argv[1]), 40, 16, 543
    rank, world, local_rank = cxdist.init_from_env()
    torch.cuda.set_device(local_rank)
    for name in ("boat_race", "hello"):
        n = n_arg if name == "boat_race" else min(n_arg, 32768)   # Hello World boards are 468 bytes each
        game = make_world(name, num_envs=n, max_episode_steps=limit, track_returns=True)
        game.its_showtime()
        acts = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        boards, rewards, discounts, flags = game.rollout_random(T, seed, env_offset=rank * n, actions_out=acts)
        stats = game.native.stats_tensor.clone()
        cxdist.all_reduce_stats(stats)                          # SUM / MAX over ranks (NCCL)
        gathered = []
        for x in (boards, rewards, flags, acts):
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x.contiguous())
            gathered.append(torch.cat(parts, dim=1))            # [T, world*n, ...] in global env order
        # a sample of THIS rank's envs through the CPU oracle
        a, b = acts.cpu().numpy(), boards.cpu().numpy()
        r, f = rewards.cpu().numpy(), flags.cpu().numpy()
        for i in (0, n // 2, n - 1):
            for t, (o, rew, dsc, term, trunc, eng) in enumerate(
                    O.rollout(name, a[:, i], rebuild_on_done=True, max_episode_steps=limit)):
                assert np.array_equal(b[t, i], np.asarray(o.board).astype(np.uint8)), (name, rank, i, t)
                assert (0.0 if rew is None else float(rew)) == float(r[t, i]), (name, rank, i, t)
                assert trunc == bool(f[t, i] & N.CX_FLAG_TRUNCATED) and term == bool(f[t, i] & N.CX_FLAG_TERMINATED)
        if rank == 0:
            whole = make_world(name, num_envs=world * n, max_episode_steps=limit, track_returns=True)
            whole.its_showtime()
            wa = torch.empty((T, world * n), dtype=torch.uint8, device="cuda")
            wb, wr, wd, wf = whole.rollout_random(T, seed, env_offset=0, actions_out=wa)
            for got, want, what in zip(gathered, (wb, wr, wf, wa), ("board", "reward", "flags", "actions")):
                assert torch.equal(got, want), "%s: %d ranks x %d envs != one GPU with %d envs (%s)" % (
                    name, world, n, world * n, what)
            one = whole.native.stats_tensor
            assert torch.equal(stats, one), (stats.tolist(), one.tolist())
        dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK world=%d envs_per_rank=%d" % (world, n_arg), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
