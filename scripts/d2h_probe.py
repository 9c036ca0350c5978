"""Development aid: host-link characterisation behind the multi-GPU end-to-end number.

Run under torchrun, one rank per GPU.  Every rank copies a 256 MB device buffer to pinned host memory (and back)
(a) alone, ranks taking turns, and (b) all ranks at once; rank 0 prints one JSON object with per-GPU and aggregate
GB/s, the NUMA node each GPU reports and the host's CPU / NUMA layout.  The end-to-end leg of bench.py moves 31.5 MB
per step and GPU device-to-host, so its ceiling is (aggregate D2H GB/s) / 30 B per env-step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/d2h_probe.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from campx_b200 import dist as cxdist


def bw(dst, src, reps, stream):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
    stream.synchronize()
    return src.numel() * src.element_size() * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, local = cxdist.init_from_env()
    torch.cuda.set_device(local)
    nbytes = 256 << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    s = torch.cuda.Stream()
    for _ in range(2):
        bw(host, dev, 2, s)
    out = {"d2h_alone": None, "h2d_alone": None}
    alone_d2h, alone_h2d = 0.0, 0.0
    for r in range(world):                       # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone_d2h, alone_h2d = bw(host, dev, 8, s), bw(dev, host, 8, s)
    if world > 1:
        dist.barrier()
    both_d2h = bw(host, dev, 16, s)              # everybody at once
    if world > 1:
        dist.barrier()
    both_h2d = bw(dev, host, 16, s)
    props = torch.cuda.get_device_properties(local)
    node = None
    try:
        with open("/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (props.pci_domain_id, props.pci_bus_id,
                                                                       props.pci_device_id)) as f:
            node = int(f.read())
    except Exception:
        pass
    mine = torch.tensor([alone_d2h, alone_h2d, both_d2h, both_h2d, -1.0 if node is None else float(node)],
                        dtype=torch.float64, device="cuda")
    parts = [torch.empty_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, mine)
    else:
        parts = [mine]
    if rank == 0:
        rows = [p.cpu().tolist() for p in parts]
        nodes = []
        try:
            nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except Exception:
            pass
        print(json.dumps({
            "gpus": world, "buffer_mb": nbytes >> 20,
            "d2h_alone_gbs": [round(r[0], 1) for r in rows], "h2d_alone_gbs": [round(r[1], 1) for r in rows],
            "d2h_concurrent_gbs": [round(r[2], 1) for r in rows], "h2d_concurrent_gbs": [round(r[3], 1) for r in rows],
            "d2h_concurrent_total_gbs": round(sum(r[2] for r in rows), 1),
            "h2d_concurrent_total_gbs": round(sum(r[3] for r in rows), 1),
            "gpu_numa_node": [int(r[4]) for r in rows], "host_numa_nodes": nodes,
            "host_cpus": len(os.sched_getaffinity(0)),
            "e2e_ceiling_env_steps_per_s": round(sum(r[2] for r in rows) * 1e9 / 30.0)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
