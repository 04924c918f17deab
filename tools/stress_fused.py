"""Stress the fused kernel: many back-to-back calls per workload, progress printed (run under `timeout`).
    python tools/stress_fused.py [calls] [timing 0/1] [workloads...]"""
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402
from mx_quantization_b200 import _lib  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 50
timing = int(sys.argv[2]) if len(sys.argv) > 2 else 0
names = sys.argv[3:] or ["deit_base_c2", "dit_xl2_c3", "pixart_c4"]
dev = torch.device("cuda:0")
lib = _lib.load()
buf = torch.zeros(320 * 32, dtype=torch.int64, device=dev)
for name in names:
    w = bench.WORKLOADS[name]
    B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
    out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
    specs = bench.mx_specs(w["bfloat"], w["flush"])
    if timing:
        lib.mxp_debug_fused_timing(c_void_p(buf.data_ptr()))
    ref = None
    for i in range(calls):
        o = mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out)
        if i % 10 == 0:
            torch.cuda.synchronize()
            cur = out.clone()
            if ref is None:
                ref = cur
            print(name, "call", i, "launches", mxq.last_launch_count(), "same as first:", bool(torch.equal(cur, ref)), flush=True)
    torch.cuda.synchronize()
    lib.mxp_debug_fused_timing(c_void_p(0))
    print(name, "done", flush=True)
