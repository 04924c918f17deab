"""One DeiT-base layer per ranking mode - the short command ncu wraps for k_predict_topk_wide.
    python tools/prof_modes.py [mode] [calls]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS["deit_base_c2"]
B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3, H, hd, device="cuda", generator=g).permute(2, 0, 3, 1, 4)
specs = bench.mx_specs(w["bfloat"], w["flush"])
for _ in range(calls):
    mxq.predict_topk(qkv[0], qkv[1], specs, w["top_k"], pred_mode=mode)
torch.cuda.synchronize()
print("ok", mode)
