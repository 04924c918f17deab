"""Generate tests/golden/modes_*.npz: the reference's OTHER top-k rankings run with the
UNMODIFIED reference functions on seeded inputs.

    python tests/golden/make_golden_modes.py          (authoring container only: needs /root/reference)

  partial_Q / partial_K   funcs.exponent_approximation(...).partial_Q() / .partial_K()
                          (funcs/exponent_based_prediction.py:274-318) then `ex_q @ ex_k^T`
                          (workloads/deit/scripts/main.py:111-118)
  MXINT4                  .MXINT4() (Sanger; funcs/exponent_based_prediction.py:179-199): both sides re-quantized
                          with elem_format="int4"
  two_step_leading_ones   .two_step_leading_ones() (EXION; funcs/exponent_based_prediction.py:96-177)
  true_ex                 .exponent_based_sign_leading_ones() of the example copy of the predictor file
                          (microxscaling/examples/deit/exponent_based_prediction.py:163-178)
  exact                   the `top_k and not approx_flag` branch: top-k of
                          mx.matmul(q, k^T) * scale  (main.py:101-102,130)
followed by the same gather / softmax / scatter_ / mx.matmul(attn, v) as every mode (main.py:124,147-152).
As in make_golden.py, torch.topk is replaced by the first k of a stable descending sort (canonical tie
rule) and the raw torch.topk indices are stored beside it.
"""
import os

import numpy as np
import torch

from make_golden import HERE, _example, exponent_approximation, make_inputs, mx_matmul, mx_specs

MODES = ("partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex", "exact")

CASES = [
    # name,              B  H  N    hd  k   bfloat flush  kind     seed
    ("modes_deit",       1, 2, 48,  64, 12, 32, False, "randn",  21),
    ("modes_dit_bf16",   1, 2, 40,  72, 10, 16, False, "randn",  22),
    ("modes_pixart",     1, 2, 33,  72,  9, 32, True,  "edges",  23),
    ("modes_deit_197",   1, 1, 197, 64, 30, 32, False, "randn",  24),
]


def predictor_operands(q, k, specs, mode):
    """true_ex (exponent_based_sign_leading_ones) exists only in the example copy of the predictor file
    (microxscaling/examples/deit/exponent_based_prediction.py:163-178); everything else comes from funcs/."""
    if mode == "true_ex":
        return _example.exponent_approximation(Q=q, K=k, mx_specs=specs).exponent_based_sign_leading_ones()
    obj = exponent_approximation(Q=q, K=k, mx_specs=specs)
    return {"partial_Q": obj.partial_Q, "partial_K": obj.partial_K, "MXINT4": obj.MXINT4,
            "two_step_leading_ones": obj.two_step_leading_ones}[mode]()


def reference_mode(q, k, v, top_k, scale, specs, mode):
    out = {}
    true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa')
    true_scores = true_scores * scale
    if mode == "exact":
        rank = true_scores
    else:
        ex_q, ex_k = predictor_operands(q, k, specs, mode)
        rank = ex_q @ ex_k.transpose(-2, -1)
    out["rank_scores"] = rank
    out["topk_idx_torch"] = torch.topk(rank, top_k, dim=-1, largest=True, sorted=True).indices
    idx = torch.sort(rank, dim=-1, descending=True, stable=True).indices[..., :top_k].contiguous()
    out["idx"] = idx
    vals = true_scores.gather(dim=-1, index=idx)
    attn = torch.zeros_like(true_scores)
    attn.scatter_(-1, idx, torch.softmax(vals, dim=-1).to(attn.dtype))
    out["out"] = mx_matmul(attn, v, mx_specs=specs, mode_config='aa')
    return out


def reference_cross_mode(q, k, v, attention_mask, top_k, scale, specs, mode):
    """PixArt cross-attention (workloads/PixArt/models/MX_transformer_block.py:791-859) in the given ranking mode:
    the additive mask goes onto the true scores (:803) and onto the predicted ones (:822); "exact" is the
    else-branch of :806 (an excluded timestep / ex_pred off): torch.topk(true_scores) (:833-834)."""
    B, H, N, _ = q.shape
    S = k.shape[2]
    out = {}
    attention_mask = attention_mask.unsqueeze(1).repeat(1, H, 1, 1)
    true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa') * scale
    attn_bias = attention_mask + torch.zeros([N, S], dtype=q.dtype)
    true_scores += attn_bias
    if mode == "exact":
        rank = true_scores
    else:
        ex_q, ex_k = predictor_operands(q, k, specs, mode)
        rank = ex_q @ ex_k.transpose(-2, -1) + attn_bias
    out["rank_scores"] = rank
    out["topk_idx_torch"] = torch.topk(rank, top_k, dim=-1, largest=True, sorted=True).indices
    idx = torch.sort(rank, dim=-1, descending=True, stable=True).indices[..., :top_k].contiguous()
    out["idx"] = idx
    vals = true_scores.gather(dim=-1, index=idx)
    attn = torch.zeros_like(true_scores)
    attn.scatter_(-1, idx, torch.softmax(vals, dim=-1))
    out["out"] = mx_matmul(attn, v, mx_specs=specs, mode_config='aa')
    return out


def cross_case():
    name, B, H, Nq, S, hd, top_k, valid, bias, bfloat, flush, seed = \
        ("modes_pixart_cross", 2, 2, 64, 40, 72, 20, (13, 27), -10000.0, 32, True, 25)
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, Nq, hd, generator=g)
    k = torch.randn(B, H, S, hd, generator=g)
    v = torch.randn(B, H, S, hd, generator=g)
    mask = torch.zeros(B, S)
    for b, n in enumerate(valid):
        mask[b, :n] = 1.0
    attention_mask = ((1.0 - mask) * bias).reshape(B, 1, S)
    arrays = {"q": q.numpy(), "k": k.numpy(), "v": v.numpy(), "key_bias": attention_mask.reshape(B, S).numpy()}
    for mode in MODES:
        ref = reference_cross_mode(q, k, v, attention_mask, top_k, 1.0 / (hd ** 0.5), mx_specs(bfloat, flush), mode)
        for key, val in ref.items():
            a = val.numpy()
            arrays[f"{mode}.{key}"] = a.astype(np.int16) if "idx" in key else a
    arrays["meta"] = np.array([B, H, Nq, S, hd, top_k, bfloat, int(flush)], dtype=np.int64)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    torch.set_num_threads(1)
    cross_case()
    for name, B, H, N, hd, top_k, bfloat, flush, kind, seed in CASES:
        specs = mx_specs(bfloat, flush)
        q, k, v = make_inputs(B, H, N, hd, seed, kind)
        arrays = {"q": q.numpy(), "k": k.numpy(), "v": v.numpy()}
        for mode in MODES:
            ref = reference_mode(q, k, v, top_k, hd ** -0.5, specs, mode)
            for key, val in ref.items():
                a = val.numpy()
                if "idx" in key:
                    a = a.astype(np.int16)
                if key == "rank_scores" and N > 64:
                    continue                      # keep the big fixture small
                arrays[f"{mode}.{key}"] = a
        arrays["meta"] = np.array([B, H, N, hd, top_k, bfloat, int(flush)], dtype=np.int64)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
