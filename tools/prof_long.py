"""One long-sequence call (C5 point) - the short command ncu wraps.
    python tools/prof_long.py [N] [ratio] [calls]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, hd, B = 16, 72, max(1, 65536 // N)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
specs = bench.mx_specs(32, False)
for _ in range(calls):
    ms = []
    mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, int(ratio * N), out=out, _kernel_ms=ms)
torch.cuda.synchronize()
print(N, "kernel ms (predict_topk, prep_v, exact_attention):", [round(x, 4) for x in ms])
