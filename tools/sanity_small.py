"""A few small calls through every kernel family (short / long predictor, both attention kernels,
codes + idx outputs) - the command compute-sanitizer wraps.
    compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (B, H, N, hd, k, bfl) in [(1, 2, 197, 64, 30, 32), (1, 1, 256, 72, 154, 16), (1, 1, 37, 96, 9, 32),
                              (1, 1, 300, 72, 75, 32), (1, 1, 640, 64, 64, 16)]:
    qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
    specs = bench.mx_specs(bfl, False)
    out, mask = mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, k, return_mask=True)
    r = mxq.predict_topk(qkv[0], qkv[1], specs, k, return_idx=True, return_codes=True)
    torch.cuda.synchronize()
    assert torch.equal(r["mask"], mask)
    print("ok", B, H, N, hd, k, bfl, float(out.abs().max()))
