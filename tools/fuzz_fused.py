#!/usr/bin/env python
"""Random-shape parity fuzz of the FUSED launch (k_fused_pruned_attention, forced with set_fused_path(2)) and of the
long-sequence pair kernel: masks bit-exact against the CPU oracle, outputs within the budget, and the fused result against the
three-kernel plan.  Head counts above 296 exercise the shared last round, head_dim 64 the compile-time instantiations.
    python tools/fuzz_fused.py [--cases 40] [--seed 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from oracle import mxint8_oracle as O  # noqa: E402
from tests.helpers import assert_out_close, make_qkv, mx_specs, out_error_budget, unpack_mask  # noqa: E402


def check(out, mask, ref, v, Nk, bfloat):
    want = O.mask_words_to_dense(O.idx_to_mask_words(ref["idx"], Nk), Nk)
    assert torch.equal(unpack_mask(mask, Nk), want), "mask mismatch"
    if bfloat == 32:
        assert_out_close(out.cpu(), ref, v, Nk, bfloat, 1e-3)
    else:       # see tools/fuzz_parity.py: bf16 ties of A1 on the true scores are invisible to the budget
        err = (out.cpu() - ref["out"]).abs().amax(-1)
        budget = out_error_budget(ref, v, Nk, bfloat, 1e-3)
        assert int((err > budget).sum()) <= max(2, int(0.01 * err.numel())), "more than 1% of rows outside the budget"
        assert float(err.max()) <= 0.05 * float(ref["out"].abs().max()), "error above 5% of max|ref|"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = torch.Generator().manual_seed(args.seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    bad = 0
    for c in range(args.cases):
        long_case = c % 5 == 4
        if long_case:
            B, H, hd = ri(1, 2), ri(1, 3), [64, 72, 80, 128][ri(0, 3)]
            N = ri(257, 900)
            top_k = ri(1, N)
        else:
            heads = [ri(64, 120), ri(297, 420), ri(121, 296)][ri(0, 2)]
            H = [1, 2, 3, 4, 6][ri(0, 4)]
            B = max(1, heads // H)
            hd = [64, 64, 32, 72, 96, 40][ri(0, 5)]
            N = [197, ri(129, 256), ri(193, 224)][ri(0, 2)]
            top_k = ri(1, N - 1) if ri(0, 3) == 0 else ri(1, max(1, int(0.35 * N)))
        bfloat = 16 if ri(0, 1) else 32
        flush = bool(ri(0, 1))
        kind = ["randn", "randn", "lognormal", "edges"][ri(0, 3)] if hd >= 64 else "randn"
        q, k, v = make_qkv(B, H, N, hd, seed=3000 + c, kind=kind)
        tag = f"case {c}: B{B} H{H} N{N} hd{hd} k{top_k} bf{bfloat} flush{int(flush)} {kind}" + (" long" if long_case else "")
        try:
            specs = mx_specs(bfloat, flush)
            ref = O.pruned_attention(q, k, v, top_k, bfloat=bfloat, flush=flush, integer_scores=True)
            res = {}
            for mode in ((1, 0) if long_case else (2, 0)):
                mxq.set_fused_path(mode)
                out, mask = mxq.pruned_attention(q.cuda(), k.cuda(), v.cuda(), specs, top_k, return_mask=True)
                n = mxq.last_launch_count()
                check(out, mask, ref, v, N, bfloat)
                res[mode] = (out.cpu(), n)
            mxq.set_fused_path(True)
            print(tag, "ok, launches", {m: r[1] for m, r in res.items()})
        except Exception as e:      # noqa: BLE001
            mxq.set_fused_path(True)
            bad += 1
            print(tag, "FAILED:", repr(e)[:300])
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
