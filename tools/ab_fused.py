"""A/B timing of one layer: three-kernel path, fused (ping-pong off / on).  python tools/ab_fused.py [workload...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402
from mx_quantization_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ["deit_base_c2", "dit_xl2_c3", "pixart_c4"]):
    w = bench.WORKLOADS[name]
    B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = [torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4) for _ in range(3)]
    out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
    specs = bench.mx_specs(w["bfloat"], w["flush"])
    res = {}
    for label, fused, pp in (("three_kernels", False, 0), ("fused", True, 0), ("fused_pingpong", True, 1)):
        mxq.set_fused_path(fused)
        lib.mxp_debug_fused_pingpong(pp)
        for _ in range(3):
            for qkv in bufs:
                mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            for qkv in bufs:
                mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out)
        e1.record()
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / (reps * len(bufs))
    mxq.set_fused_path(True)
    lib.mxp_debug_fused_pingpong(1)
    print(name, {k: round(v, 4) for k, v in res.items()}, "ms per layer", flush=True)
