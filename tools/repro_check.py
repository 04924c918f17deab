import sys, time, torch
sys.path.insert(0, '.')
import mx_quantization_b200 as mxq
from bench import mx_specs
from oracle import mxint8_oracle as O
specs = mx_specs(32, False)
N, H, hd, B = 512, 16, 72, 128
qkv = torch.randn(B, N, 3, H, hd, device='cuda').permute(2, 0, 3, 1, 4)
q, k, v = qkv[0], qkv[1], qkv[2]
top_k = 52
print('start', flush=True)
o2, mask = mxq.pruned_attention(q[:1, :1], k[:1, :1], v[:1, :1], specs, top_k, return_mask=True)
torch.cuda.synchronize(); print('gpu slice ok', flush=True)
t0 = time.time()
qs, ks, vs = q[:1, :1].cpu(), k[:1, :1].cpu(), v[:1, :1].cpu()
print('copied', qs.shape, qs.stride(), flush=True)
ref = O.pruned_attention(qs, ks, vs, top_k, integer_scores=True)
print('oracle ok', time.time() - t0, flush=True)
