import sys, time, torch
sys.path.insert(0, '.')
import mx_quantization_b200 as mxq
from bench import mx_specs
specs = mx_specs(32, False)
B, N, H, hd = int(sys.argv[1]), int(sys.argv[2]), 16, 72
reps = int(sys.argv[3])
q, k, v = (torch.randn(B, H, N, hd, device='cuda') for _ in range(3))
ref = None
for it in range(reps):
    o, m = mxq.pruned_attention(q, k, v, specs, N // 4, return_mask=True)
    torch.cuda.synchronize()
    if ref is None:
        ref = (o.clone(), m.clone())
    else:
        assert torch.equal(o, ref[0]) and torch.equal(m, ref[1]), f"nondeterministic at rep {it}"
    print(it, end=' ', flush=True)
print('stress ok', B, N, flush=True)
