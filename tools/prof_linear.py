"""One MX Linear call on the DeiT-base qkv shape - the short command ncu wraps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mx_quantization_b200 as mxq
from bench import mx_specs
dev = torch.device("cuda:0")
M, K, N = 256 * 197, 768, 2304
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5; b = torch.randn(N, device=dev) * 0.1
specs = mx_specs(32, False)
w_op = mxq.mx_linear_prepare_weight(w, specs)
for _ in range(3):
    y = mxq.mx_linear(x, w_op, b, specs, out_features=N)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
