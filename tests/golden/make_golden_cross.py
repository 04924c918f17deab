"""Generate tests/golden/pixart_cross*.npz: PixArt-alpha CROSS-attention with the additive text mask
(SURVEY 8f1), by running the UNMODIFIED reference functions in the order of
workloads/PixArt/models/MX_transformer_block.py:791-859:

    true_scores = matmul(q, k^T) * scale_factor ; true_scores += attn_bias          :794-803
    ex_q, ex_k  = exponent_approximation(Q=q, K=k).exponent_based_sign()            :806-809
    pred_scores = ex_q @ ex_k^T + attn_bias                                         :821-822
    idx         = topk(pred_scores, k)   (canonical: first k of a stable descending sort)
    vals        = true_scores.gather(idx) ; softmax ; scatter_ ; matmul(attn, v)     :828-857

attn_bias is the (B,1,1,S) additive mask diffusers builds, (1 - mask) * -10000, repeated over heads
and broadcast over the query rows (:776, :796-802).  Run in the authoring container only:

    python tests/golden/make_golden_cross.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import (exponent_approximation, mx_matmul, mx_specs,  # noqa: E402
                         working_exponent_based_sign)


def reference_cross_layer(q, k, v, attention_mask, top_k, scale, specs):
    B, H, N, _ = q.shape
    S = k.shape[2]
    out = {}
    attention_mask = attention_mask.unsqueeze(1).repeat(1, H, 1, 1)              # :776  (B,H,1,S)
    true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa') * scale
    attn_bias = torch.zeros([N, S], dtype=q.dtype)
    attn_bias = attention_mask + attn_bias                                       # :800-802 -> (B,H,N,S)
    true_scores += attn_bias
    obj = exponent_approximation(Q=q, K=k, mx_specs=specs)
    ex_q, ex_k = working_exponent_based_sign(obj)
    pred = ex_q @ ex_k.transpose(-2, -1)
    pred = pred + attn_bias
    out["pred_scores"] = pred
    out["topk_idx_torch"] = torch.topk(pred, top_k, dim=-1, largest=True, sorted=True).indices
    idx = torch.sort(pred, dim=-1, descending=True, stable=True).indices[..., :top_k].contiguous()
    out["idx"] = idx
    vals = true_scores.gather(dim=-1, index=idx)
    out["true_vals"] = vals
    attn = torch.zeros_like(true_scores)
    attn.scatter_(-1, idx, torch.softmax(vals, dim=-1))
    out["out"] = mx_matmul(attn, v, mx_specs=specs, mode_config='aa')
    return out


CASES = [
    # name,             B  H  Nq   S    hd  k   valid text tokens per batch element, bias,  bfloat flush seed
    ("pixart_cross",     2, 2, 64,  40, 72, 20, (13, 27),                            -10000.0, 32, True, 21),
    ("pixart_cross_k77", 1, 2, 96, 120, 72, 77, (31,),                               -10000.0, 32, True, 22),
    ("pixart_cross_all", 1, 2, 48,  40, 64, 12, (40,),                               -10000.0, 16, False, 23),
]


def main():
    torch.set_num_threads(1)
    for name, B, H, Nq, S, hd, top_k, valid, bias, bfloat, flush, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        q = torch.randn(B, H, Nq, hd, generator=g)
        k = torch.randn(B, H, S, hd, generator=g)
        v = torch.randn(B, H, S, hd, generator=g)
        mask = torch.zeros(B, S)
        for b, n in enumerate(valid):
            mask[b, :n] = 1.0
        attention_mask = ((1.0 - mask) * bias).reshape(B, 1, S)                  # diffusers: (B,1,S) -> unsqueeze(1)
        scale = 1.0 / (hd ** 0.5)
        ref = reference_cross_layer(q, k, v, attention_mask, top_k, scale, mx_specs(bfloat, flush))
        arrays = {"q": q, "k": k, "v": v, "key_bias": attention_mask.reshape(B, S), **ref}
        np_arrays = {n_: (a.numpy() if isinstance(a, torch.Tensor) else a) for n_, a in arrays.items()}
        np_arrays["meta"] = np.array([B, H, Nq, S, hd, top_k, bfloat, int(flush)], dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **np_arrays)
        print(name, {n_: a.shape for n_, a in np_arrays.items()})


if __name__ == "__main__":
    main()
