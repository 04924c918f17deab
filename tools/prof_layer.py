"""One layer of a bench workload, a few calls - the short command ncu wraps.
    python tools/prof_layer.py [workload] [calls] [three]      (three: the three-kernel path, mxp_set_fused_path(0))"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "deit_base_c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
specs = bench.mx_specs(w["bfloat"], w["flush"])
if len(sys.argv) > 3 and sys.argv[3] == "three":
    mxq.set_fused_path(False)
for _ in range(calls):
    ms = []
    mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out, _kernel_ms=ms)
torch.cuda.synchronize()
print(name, "kernel ms (predict_topk, prep_v, exact_attention):", [round(x, 4) for x in ms])
