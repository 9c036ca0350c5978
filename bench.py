#!/usr/bin/env python
"""Benchmark of the batched Engine.play() hot path: boat_race, 2^20 environments per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one Engine.play() over the whole environment batch (N_envs env-steps per GPU).  The K timed
steps are issued as fused `cx_rollout` launches of `--chunk` steps each (state stays on chip inside a
launch); inputs (uint8 action indices, generated beforehand with the library's Philox kernel) are
resident in HBM when the timed region starts, and every env-step writes its full observation contract
(board 25 B, reward f32, flags u8) to HBM.  One launch writes ~1 GB, far more than the 126 MB L2, and
consecutive launches alternate between two output buffers.

JSON line keys follow the driver contract; see DESIGN.md "Measurement" for how each number is taken.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORLD = "boat_race"
EPISODE_LIMIT = 100          # examples/actor_critic.py:56
SEED = 543                   # examples/actor_critic.py:26


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16384)
    ap.add_argument("--warmup", type=int, default=512)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="environments per GPU")
    ap.add_argument("--chunk", type=int, default=32, help="env-batch steps fused per kernel launch")
    ap.add_argument("--e2e-steps", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--obs-reps", type=int, default=8,
                    help="launches of the secondary board+layered-board measurement (0: skip it)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Engine.play (numpy), one env per object, all host cores
# --------------------------------------------------------------------------------------------------

def _cpu_worker(job):
    n_envs, warm, steps, seed = job
    import numpy as np
    from oracle import campx_oracle as O
    rng = np.random.Generator(np.random.PCG64(seed))
    envs = [O.World(WORLD) for _ in range(n_envs)]
    age = [0] * n_envs

    def batch_step():
        acts = rng.integers(0, 5, size=n_envs)
        for i in range(n_envs):
            envs[i].step(int(acts[i]))
            age[i] += 1
            if age[i] >= EPISODE_LIMIT:          # a fresh make_game() per episode (actor_critic.py:146)
                envs[i] = O.World(WORLD)
                age[i] = 0

    for _ in range(warm):
        batch_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        batch_step()
    return n_envs * steps, time.perf_counter() - t0


def cpu_reference_rate(steps, warm, budget_s, cores=None):
    """env-steps/s of the oracle port on `cores` host processes; each of `steps` steps advances a
    bounded sample of environments (sized from a calibration run to fit `budget_s` seconds)."""
    import multiprocessing as mp
    if cores is None:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    done, dt = _cpu_worker((4, 2, 40, 1))                      # calibration: one core
    rate1 = done / dt
    per_core = int(max(1, min(256, rate1 * budget_s / max(1, steps + warm))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(per_core, warm, steps, SEED + i) for i in range(cores)])
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    sample = "%d procs x %d envs x %d steps of %s, episode rebuilt every %d steps, oracle/campx_oracle.py (numpy port)" % (
        cores, per_core, steps, WORLD, EPISODE_LIMIT)
    return total / slowest, cores, sample, slowest


def workload_config(envs, world, chunk):
    """`config` of the JSON line; identical for both arms (the reference arm times a bounded sample of it)."""
    per_step = 1 + 4 + 1 + 25
    return {"workload": "boat_race 5x5 (examples/worlds.py == reference examples/boat_race.py), "
                        "2^20 envs per GPU, uniform random actions, episode limit 100 + auto reset",
            "envs_per_gpu": envs, "total_envs": world * envs, "fused_steps_per_launch": chunk,
            "observation_contract": "board u8[25] + reward f32 + flags u8 per env-step",
            "l2_policy": "outputs larger than L2: each launch writes %.0f MB, two alternating output "
                         "buffers" % (envs * chunk * (per_step - 1) / 1e6),
            "parallelism": "env-sharded x%d, no step-path collective" % world}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2000))
    warm = max(0, min(args.warmup, 20))
    value, cores, sample, elapsed = cpu_reference_rate(steps, warm, budget_s=60.0)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": elapsed * 1e3 / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.envs, max(1, args.gpus), max(1, args.chunk)),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the benchmark runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = False          # only samples taken while `active` count as "under load"
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.active:
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------

def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(key):
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from campx_b200 import dist as cxdist
    from examples.worlds import make_world

    rank, world, local_rank = cxdist.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n, T, K, W = args.envs, max(1, args.chunk), args.steps, args.warmup

    sampler = ClockSampler(local_rank)
    sampler.start()

    game = make_world(WORLD, num_envs=n, max_episode_steps=EPISODE_LIMIT, auto_reset=True, track_returns=True)
    game.its_showtime()                      # compiles the user-level world and uploads it
    nat = game.native
    env_offset = rank * n                    # rank r owns envs [r*n, (r+1)*n): independent Philox streams
    n_act = 4
    actions = [nat.fill_actions(T, seed=SEED, env_offset=env_offset, t0=i * T) for i in range(n_act)]
    outs = [game.alloc_rollout(T) for _ in range(2)]

    def run_steps(k, counter):
        """k env-batch steps as fused launches of up to T steps; returns launches issued."""
        launches, i = 0, counter
        while k > 0:
            t = min(T, k)
            a = actions[i % n_act]
            o = outs[i % 2]
            if t == T:
                game.rollout(a, o)
            else:
                game.rollout(a[:t], tuple(None if x is None else x[:t] for x in o))
            k -= t
            launches += 1
            i += 1
        return launches

    def barrier():
        if world > 1:
            dist.barrier()

    run_steps(max(W, 3), 0)                  # warm-up (>= 3 steps)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    torch.cuda.synchronize()
    e0.record()
    launches = run_steps(K, 1)
    e1.record()
    torch.cuda.synchronize()
    sampler.active = False
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms.item())
    value = world * n * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_agent_rollout): algorithmic bytes / avg launch duration ----
    cells = nat.cells
    per_step = 1 + 4 + 1 + cells                     # action u8 + reward f32 + flags u8 + board u8[25]
    state_rw = 2 * (1 + 2 + 4)                       # per launch and env: cell u8, step u16, return f32 (r+w)
    bytes_per_launch_full = n * (T * per_step + state_rw)
    alg_bytes_total = n * (K * per_step + launches * state_rw)
    achieved = alg_bytes_total / (ms * 1e-3) / 1e9   # this rank's kernel stream; ms is the max over ranks
    peak, peak_src = measured_peak()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("k_agent_rollout_track_T%d_n%d" % (T, n)),
                "kernel": "k_agent_rollout<TRACK=true>", "alg_bytes_per_env_step": per_step + state_rw / T,
                "alg_bytes_per_launch": bytes_per_launch_full, "avg_launch_ms": ms / launches,
                "peak_source": peak_src}

    # ---- end to end through Engine.play() with HOST buffers (H2D actions, D2H board+reward+flags) ----
    # Every step: pinned-host actions -> device, one Engine.play(), the whole result (board, reward, flags) ->
    # pinned host.  play() alternates between two output buffer sets, so the D2H of step t runs on a side
    # stream while the H2D + kernel of step t+1 run on the main stream (PCIe is full duplex); events keep a
    # buffer set from being overwritten before its copy-out has finished.
    E = max(1, args.e2e_steps)
    h_act = torch.randint(0, 5, (E, n), dtype=torch.uint8).pin_memory()
    h_out = [(torch.empty((n, 5, 5), dtype=torch.uint8).pin_memory(), torch.empty((n,), dtype=torch.float32).pin_memory(),
              torch.empty((n,), dtype=torch.uint8).pin_memory()) for _ in range(2)]
    d_act = [torch.empty((n,), dtype=torch.uint8, device=dev) for _ in range(2)]
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev)
    stepped = [torch.cuda.Event() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    for ev in copied:
        ev.record(main)

    def e2e_step(i):
        k = i % 2
        main.wait_event(copied[k])                   # the buffers this play() will overwrite have been copied out
        d_act[k].copy_(h_act[i % E], non_blocking=True)
        obs, reward, _ = game.play(d_act[k])
        flags = game.flags
        stepped[k].record(main)
        with torch.cuda.stream(side):
            side.wait_event(stepped[k])
            h_out[k][0].copy_(obs.board, non_blocking=True)
            h_out[k][1].copy_(reward, non_blocking=True)
            h_out[k][2].copy_(flags, non_blocking=True)
            copied[k].record(side)

    for i in range(4):
        e2e_step(i)
    torch.cuda.synchronize()
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for i in range(E):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.active = False
    if world > 1:
        t_s = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        e2e_s = float(t_s.item())
    e2e = {"value": world * n * E / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": n,
           "d2h_bytes_per_step": n * (cells + 4 + 1), "steps": E,
           "what": "Engine.play(): pinned-host uint8 actions -> H2D -> cx_step -> D2H of board, reward, flags every step "
                   "(copy-out of step t on a side stream overlaps H2D + kernel of step t+1)"}

    # ---- secondary contract (BASELINE config 5's policy input): board + layered board per env-step ----
    obs_extra = None
    try:
        if args.obs_reps <= 0:
            raise RuntimeError("skipped (--obs-reps 0)")
        To = 16
        del outs
        torch.cuda.empty_cache()
        o_outs = [game.alloc_rollout(To) for _ in range(2)]
        o_lay = [torch.empty((To, n, nat.n_chars, 5, 5), dtype=torch.uint8, device=dev) for _ in range(2)]
        for i in range(3):
            game.rollout_observations(actions[i % n_act][:To], o_outs[i % 2], o_lay[i % 2])
        torch.cuda.synchronize()
        reps = args.obs_reps
        e0.record()
        for i in range(reps):
            game.rollout_observations(actions[i % n_act][:To], o_outs[i % 2], o_lay[i % 2])
        e1.record()
        torch.cuda.synchronize()
        oms = e0.elapsed_time(e1) / reps
        ob = 1 + 4 + 1 + cells * (1 + nat.n_chars)
        obs_extra = {"kernel": "k_agent_rollout_obs<TRACK=true>", "alg_bytes_per_env_step": ob,
                     "env_steps_per_sec": n * To / (oms * 1e-3), "achieved_gbs": n * To * ob / (oms * 1e-3) / 1e9,
                     "frac_of_measured_peak": n * To * ob / (oms * 1e-3) / 1e9 / peak, "avg_launch_ms": oms,
                     "fused_steps_per_launch": To, "what": "board u8[25] + layered board u8[7,25] + reward + flags"}
        del o_outs, o_lay
    except Exception as exc:                                     # secondary number: never fail the bench line
        obs_extra = {"error": repr(exc)}

    # ---- episode-return statistics: the one collective of the design (not on the step path) ----
    stats = nat.stats_tensor.clone()
    cxdist.all_reduce_stats(stats)
    summary = cxdist.summarize_stats(stats.cpu())
    clocks = sampler.stop()

    if rank == 0:
        cpu = None
        if world == 1:
            v, cores, sample, _ = cpu_reference_rate(steps=200, warm=5, budget_s=args.cpu_seconds)
            cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(n, world, T),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "return_stats": summary, "observation_contract_kernel": obs_extra,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the run, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    # stdout carries ONE JSON line: NCCL prints its version banner there (NCCL_DEBUG=VERSION/INFO, from the
    # environment or an nccl.conf), so everything but the JSON line is sent to stderr at the descriptor level
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "INFO"):
        os.environ["NCCL_DEBUG"] = "WARN"
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
