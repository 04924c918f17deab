#!/usr/bin/env python
"""Bare pinned-host <-> device copy bandwidth with N ranks copying at the same time: the ceiling of bench.py's
`e2e` figure.  Every rank owns one pinned buffer and one device buffer of --mb MiB and runs
--reps copies per direction between two barriers; prints per-rank and aggregate GB/s, and the CPU affinity /
NUMA node the rank ran on.
    python tools/h2d_scaling.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_scaling.py"""
import argparse
import json
import os

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    gpu = torch.empty(n, dtype=torch.uint8, device=dev)
    res = {}
    # both directions at once, in bench.py's e2e proportion (three input buffers in for one output buffer out)
    host2 = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
    gpu2 = torch.empty(n // 3, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_in.wait_stream(torch.cuda.current_stream())
    s_out.wait_stream(torch.cuda.current_stream())
    for _ in range(args.reps):
        with torch.cuda.stream(s_in):
            gpu.copy_(host, non_blocking=True)
        with torch.cuda.stream(s_out):
            host2.copy_(gpu2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_in)
    torch.cuda.current_stream().wait_stream(s_out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["both_3to1"] = {"rank_gbs": (n + n // 3) * args.reps / (ms * 1e-3) / 1e9,
                        "aggregate_gbs": world * (n + n // 3) * args.reps / (float(t.item()) * 1e-3) / 1e9}
    for name, (dst, src) in (("h2d", (gpu, host)), ("d2h", (host, gpu))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"rank_gbs": n * args.reps / (ms * 1e-3) / 1e9,
                     "aggregate_gbs": world * n * args.reps / (float(t.item()) * 1e-3) / 1e9}
    try:
        aff = sorted(os.sched_getaffinity(0))
        aff = f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)"
    except Exception:
        aff = "?"
    numa = "?"
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        p = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:00.0/numa_node"
        numa = open(p).read().strip()
    except Exception:
        pass
    line = {"n_ranks": world, "rank": rank, "mb": args.mb, "reps": args.reps, "gpu_numa_node": numa, "cpu_affinity": aff, **res}
    if dist:
        out = [None] * world
        dist.all_gather_object(out, line)
        if rank == 0:
            print(json.dumps({"n_ranks": world, "h2d_aggregate_gbs": res["h2d"]["aggregate_gbs"],
                              "d2h_aggregate_gbs": res["d2h"]["aggregate_gbs"],
                              "both_3to1_aggregate_gbs": res["both_3to1"]["aggregate_gbs"], "ranks": out}), flush=True)
        dist.destroy_process_group()
    else:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
