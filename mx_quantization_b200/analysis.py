"""Analysis outputs of the reference's --anal mode that follow directly from the kept-key bitmask
(SURVEY.md 8 f4).  Pure torch bit operations on the mask the kernels already produce - they run on
whatever device the mask lives on and need no extra kernel.

    coverage_rate      funcs/analysis.py:56-110  total_chosen_k: distinct keys chosen by any row of a
                       head, divided by the number of query rows, averaged over batch x heads
    topk_overlap       fraction of a reference index set (e.g. torch.topk of the true scores,
                       workloads/deit/scripts/main.py:130) that the predictor kept, per row
    diff_idx_analysis  funcs/analysis.py:136-157, the overlap score the --anal text files hold, computed as the
                       reference computes it (pinned by tests/golden/analysis_overlap.npz)
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
    rows = m.shape[-2]
    while rows > 1:                              # OR over the query rows by halving (bitwise_or has no reduce): log2(Nq) ops
        half = rows // 2
        lo = m[..., :half, :] | m[..., half:2 * half, :]
        m = torch.cat([lo, m[..., 2 * half:rows, :]], dim=-2) if rows & 1 else lo
        rows = m.shape[-2]
    unique = _popcount32(m[..., 0, :]).sum(-1).to(torch.float64)           # (B,H) distinct keys
    return float((unique / mask.shape[-2]).mean())


def topk_overlap(mask: torch.Tensor, ref_idx: torch.Tensor) -> torch.Tensor:
    """Per row: |kept ∩ ref_idx| / |ref_idx|.  ref_idx int64 (B,H,Nq,k)."""
    m = mask.to(torch.int64) & 0xFFFFFFFF
    words = torch.gather(m, -1, ref_idx >> 5)
    hit = (words >> (ref_idx & 31)) & 1
    return hit.to(torch.float64).mean(-1)


def diff_idx_analysis(true_idx: torch.Tensor, pred_idx: torch.Tensor) -> float:
    """The reference's --anal overlap score, as it computes it (funcs/analysis.py:136-157; callers: the `anal` branches of
    workloads/DiT/models.py and workloads/deit/scripts/main.py): per row, the sum of the entries of ``true_idx`` that occur
    in ``pred_idx`` divided by the sum of all its entries; summed over the first 100 batch entries and divided by
    100 * heads * rows.  ``torch.isin`` tests membership in the WHOLE ``pred_idx`` tensor (not row by row), and the
    normalisation is by 100 whatever the batch size - both kept as the reference has them.  Use :func:`topk_overlap` for
    the per-row intersection ratio.  true_idx, pred_idx: integer (B,H,Nq,k) index tensors."""
    present = torch.isin(true_idx, pred_idx)
    kept = torch.where(present, true_idx, torch.zeros_like(true_idx))
    ratio = kept.sum(dim=-1, keepdim=True) / true_idx.sum(dim=-1, keepdim=True)
    total = ratio[0:100, :, :, 0].sum().item()
    return total / (100 * ratio.shape[1] * ratio.shape[2])
