"""Generate tests/golden/elsa_*.npz: the reference's ELSA ranking (funcs/elsa_approximation.py, called from
workloads/deit/scripts/main.py:119-124) run with the UNMODIFIED reference class on seeded inputs and a projection
matrix from the reference's own generator (_create_structured_orthogonal_matrix: Kronecker products of
Gram-Schmidt bases, 64 = 4x4x4, 72 = 8x9).

    python tests/golden/make_golden_elsa.py          (authoring container only: needs /root/reference)
"""
import os

import numpy as np
import torch

from make_golden import HERE, make_inputs, mx_matmul, mx_specs
from funcs.elsa_approximation import _create_structured_orthogonal_matrix, elsa_approximation

CASES = [
    # name,        B  H  N    hd  k   bfloat flush  kind     seed
    ("elsa_deit",  1, 2, 197, 64, 30, 32, False, "randn",  41),
    ("elsa_dit",   1, 2, 96,  72, 40, 16, False, "randn",  42),
    ("elsa_edges", 1, 2, 40,  64, 10, 32, False, "edges",  43),
]


def main():
    torch.set_num_threads(1)
    for name, B, H, N, hd, top_k, bfloat, flush, kind, seed in CASES:
        specs = mx_specs(bfloat, flush)
        q, k, v = make_inputs(B, H, N, hd, seed, kind)
        torch.manual_seed(seed)
        P = _create_structured_orthogonal_matrix(hd)
        true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa') * (hd ** -0.5)
        obj = elsa_approximation(Q=q, K=k, mx_specs=specs, orthogonal_matrix=P)
        rank = obj.approximation_scores()
        # smallest |projection| relative to the row norm: how far every hash bit is from flipping
        mq, mk = obj.MX_Q.double(), obj.MX_K.double()
        margin = min(float(((m @ P.double().T).abs() / m.norm(dim=-1, keepdim=True).clamp(min=1e-30)).min())
                     for m in (mq[mq.abs().sum(-1) > 0].unsqueeze(0), mk[mk.abs().sum(-1) > 0].unsqueeze(0)))
        idx = torch.sort(rank, dim=-1, descending=True, stable=True).indices[..., :top_k].contiguous()
        vals = true_scores.gather(dim=-1, index=idx)
        attn = torch.zeros_like(true_scores)
        attn.scatter_(-1, idx, torch.softmax(vals, dim=-1))
        out = mx_matmul(attn, v, mx_specs=specs, mode_config='aa')
        arrays = {"q": q.numpy(), "k": k.numpy(), "v": v.numpy(), "P": P.numpy(), "rank_scores": rank.numpy(),
                  "idx": idx.numpy().astype(np.int16), "out": out.numpy(),
                  "topk_idx_torch": torch.topk(rank, top_k, dim=-1).indices.numpy().astype(np.int16),
                  "hash_margin": np.array([margin]),
                  "meta": np.array([B, H, N, hd, top_k, bfloat, int(flush)], dtype=np.int64)}
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB, hash margin {margin:.2e}")


if __name__ == "__main__":
    main()
