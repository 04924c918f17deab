"""Per-layer time of the selection kernel and of the whole pruned attention for each ranking mode
(ex_pred / partial_Q / partial_K / exact) on the DeiT-base and DiT-XL/2 layer shapes.
    python tools/bench_modes.py [--reps 10] > profiles/rNN_ranking_modes_b200.jsonl
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mx_quantization_b200 as mxq  # noqa: E402
from tests.helpers import mx_specs  # noqa: E402

SHAPES = {"deit_base": (256, 12, 197, 64, 30, 32), "dit_xl2": (256, 16, 256, 72, 154, 16)}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    for name, (B, H, N, hd, k, bfloat) in SHAPES.items():
        specs = mx_specs(bfloat, False)
        g = torch.Generator(device="cuda").manual_seed(0)
        q, kk, v = (torch.randn(B, H, N, hd, device="cuda", generator=g) for _ in range(3))
        P = torch.linalg.qr(torch.randn(hd, hd, generator=torch.Generator().manual_seed(1)))[0].cuda()
        for mode in ("ex_pred", "partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex", "ELSA", "exact"):
            extra = {"orthogonal_matrix": P} if mode == "ELSA" else {}
            t_sel = timed(lambda: mxq.predict_topk(q, kk, specs, k, pred_mode=mode, **extra), args.reps)
            t_all = timed(lambda: mxq.pruned_attention(q, kk, v, specs, k, pred_mode=mode, **extra), args.reps)
            print(json.dumps({"workload": name, "pred_mode": mode, "B": B, "H": H, "N": N, "hd": hd, "top_k": k,
                              "select_ms": round(t_sel, 4), "layer_ms": round(t_all, 4),
                              "heads_per_s": round(B * H / (t_all * 1e-3))}), flush=True)


if __name__ == "__main__":
    main()
