import sys, time, torch
sys.path.insert(0, '.')
import mx_quantization_b200 as mxq
from bench import mx_specs
specs = mx_specs(32, False)
B, N, H, hd = int(sys.argv[1]), int(sys.argv[2]), 16, 72
which = sys.argv[3]
q, k, v = (torch.randn(B, H, N, hd, device='cuda') for _ in range(3))
top_k = N // 4
t0 = time.time()
if which == 'k1':
    r = mxq.predict_topk(q, k, specs, top_k)
elif which == 'k2':
    r = mxq.predict_topk(q[:1, :1].expand(1, 1, N, hd).contiguous(), k[:1, :1].contiguous(), specs, top_k, return_codes=True)
    mask = r["mask"].expand(B, H, N, -1).contiguous()
    qc, qe = mxq.quantize_mxint8(q, specs); kc, ke = mxq.quantize_mxint8(k, specs)
    torch.cuda.synchronize(); t0 = time.time()
    o = mxq.sparse_attention(qc, qe, kc, ke, v, mask, specs)
else:
    o = mxq.pruned_attention(q, k, v, specs, top_k)
torch.cuda.synchronize()
print(which, B, N, 'ok', round((time.time() - t0) * 1e3, 2), 'ms', flush=True)
