#!/usr/bin/env python
"""Host-side cost of one pruned_attention call (Python mirror + ctypes + tensor maps + launch) against the kernel time,
on a call small enough to be launch-bound, eager and replayed from a CUDA graph.
    python tools/host_overhead.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

dev = torch.device("cuda:0")
specs = bench.mx_specs(32, False)
for B in (6, 32, 256):
    H, N, hd, k = 12, 197, 64, 30
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
    out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
    L = 12

    def step():
        for _ in range(L):
            mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, k, out=out)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / reps
    # the same step captured once and replayed
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            step()
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"B": B, "heads": B * H, "layers": L, "eager_ms_per_step": eager_ms, "host_issue_ms_per_step": 1e3 * t_issue / reps,
                      "host_us_per_call": 1e6 * t_issue / reps / L, "graph_ms_per_step": graph_ms}), flush=True)
