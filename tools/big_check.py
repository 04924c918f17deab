import sys, torch
sys.path.insert(0, '/root/repo')
import mx_quantization_b200 as mxq
from oracle import mxint8_oracle as O
from tests.helpers import assert_out_close, mx_specs, unpack_mask
for (B, H, N, hd, k, bf) in [(1, 1, 8192, 64, 800, 32), (1, 2, 5000, 72, 1250, 16)]:
    g = torch.Generator().manual_seed(5)
    q, kk, v = (torch.randn(B, H, N, hd, generator=g) for _ in range(3))
    out, mask = mxq.pruned_attention(q.cuda(), kk.cuda(), v.cuda(), mx_specs(bf, False), k, return_mask=True)
    torch.cuda.synchronize()
    ref = O.pruned_attention(q, kk, v, k, bfloat=bf, integer_scores=True)
    want = O.mask_words_to_dense(O.idx_to_mask_words(ref["idx"], N), N)
    print(N, hd, "mask equal:", bool(torch.equal(unpack_mask(mask, N), want)), "max err rel:", float((out.cpu() - ref["out"]).abs().max() / ref["out"].abs().max()))
