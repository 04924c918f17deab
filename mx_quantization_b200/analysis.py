"""Analysis outputs of the reference's --anal mode that follow directly from the kept-key bitmask
(SURVEY.md 8 f4).  Pure torch bit operations on the mask the kernels already produce - they run on
whatever device the mask lives on and need no extra kernel.

    coverage_rate      funcs/analysis.py:56-110  total_chosen_k: distinct keys chosen by any row of a
                       head, divided by the number of query rows, averaged over batch x heads
    topk_overlap       fraction of a reference index set (e.g. torch.topk of the true scores,
                       workloads/deit/scripts/main.py:130) that the predictor kept, per row
"""
import torch


def _popcount32(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.int64) & 0xFFFFFFFF
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return (x * 0x01010101 >> 24) & 0xFF


def coverage_rate(mask: torch.Tensor) -> float:
    """mask: int32 (B,H,Nq,ceil(Nk/32)) from pruned_attention(..., return_mask=True) / predict_topk."""
    m = mask.to(torch.int64) & 0xFFFFFFFF
    union = m[..., 0, :].clone()
    for r in range(1, m.shape[-2]):             # OR over the query rows (bitwise_or has no reduce)
        union |= m[..., r, :]
    unique = _popcount32(union).sum(-1).to(torch.float64)           # (B,H) distinct keys
    return float((unique / mask.shape[-2]).mean())


def topk_overlap(mask: torch.Tensor, ref_idx: torch.Tensor) -> torch.Tensor:
    """Per row: |kept ∩ ref_idx| / |ref_idx|.  ref_idx int64 (B,H,Nq,k)."""
    m = mask.to(torch.int64) & 0xFFFFFFFF
    words = torch.gather(m, -1, ref_idx >> 5)
    hit = (words >> (ref_idx & 31)) & 1
    return hit.to(torch.float64).mean(-1)
