#!/usr/bin/env python
"""Random-shape parity fuzz: pruned_attention (masks bit-exact, outputs within the budget) and mx_linear
against the CPU oracle over random (B, H, Nq, Nk, head_dim, top_k, bfloat, flush, input kind, bias).
    python tools/fuzz_parity.py [--cases 60] [--seed 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from oracle import mxint8_oracle as O  # noqa: E402
from tests.helpers import assert_out_close, make_qkv, mx_specs, unpack_mask  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = torch.Generator().manual_seed(args.seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    bad = 0
    for c in range(args.cases):
        B, H = ri(1, 2), ri(1, 3)
        hd = [32, 40, 48, 64, 72, 80, 96, 104, 128, 36, 20][ri(0, 10)]
        long_case = ri(0, 5) == 0
        Nk = ri(257, 700) if long_case else ri(1, 256)
        Nq = Nk if ri(0, 2) else ri(1, 300)
        top_k = ri(1, Nk)
        bfloat = 16 if ri(0, 1) else 32
        flush = bool(ri(0, 1))
        kind = ["randn", "lognormal", "edges"][ri(0, 2)] if min(Nq, Nk) >= 16 and Nq == Nk and hd >= 64 else "randn"
        biased = (not long_case) and hd % 8 == 0 and ri(0, 3) == 0
        q, _, _ = make_qkv(B, H, Nq, hd, seed=1000 + c, kind=kind)
        _, k, v = make_qkv(B, H, Nk, hd, seed=2000 + c, kind=kind)
        bias = None
        if biased:
            keep = (torch.rand(B, Nk, generator=rng) < 0.4).float()
            bias = ((1 - keep) * (-10000.0 if ri(0, 1) else -3.25)).reshape(B, 1, 1, Nk)
        tag = f"case {c}: B{B} H{H} Nq{Nq} Nk{Nk} hd{hd} k{top_k} bf{bfloat} flush{int(flush)} {kind} bias{int(biased)}"
        try:
            specs = mx_specs(bfloat, flush)
            if hd % 8 and (biased or True):
                mxq.set_attention_path("cuda_core") if Nk <= 256 else None
            if hd % 8 and Nk > 256:
                print(tag, "skipped (tensor-core attention needs head_dim % 8 == 0 for Nk > 256)")
                continue
            out, mask = mxq.pruned_attention(q.cuda(), k.cuda(), v.cuda(), specs, top_k, return_mask=True,
                                             key_bias=None if bias is None else bias.cuda())
            mxq.set_attention_path("tcgen05")
            ref = O.pruned_attention(q, k, v, top_k, bfloat=bfloat, flush=flush, integer_scores=True, key_bias=bias)
            want = O.mask_words_to_dense(O.idx_to_mask_words(ref["idx"], Nk), Nk)
            got = unpack_mask(mask, Nk)
            assert torch.equal(got, want), "mask mismatch"
            if bfloat == 32:
                assert_out_close(out.cpu(), ref, v, Nk, bfloat, 1e-3)
            else:
                # bfloat 16 adds two more discontinuities the budget cannot see from the outputs alone:
                # A1 on the true scores (a one-ulp difference in an fp32 sum that is not exactly
                # representable - wide exponent spreads - flips a bf16 tie and moves that logit by up to
                # 2^-8 relative) and A1 on the output.  Masks stay bit-exact; for the outputs require
                # 99% of the rows inside the budget and no error above 5% of max|ref|.
                from tests.helpers import out_error_budget
                err = (out.cpu() - ref["out"]).abs().amax(-1)
                budget = out_error_budget(ref, v, Nk, bfloat, 1e-3)
                nbad = int((err > budget).sum())
                assert nbad <= max(2, int(0.01 * err.numel())), "more than 1% of rows (and more than 2) outside the budget"
                assert float(err.max()) <= 0.05 * float(ref["out"].abs().max()), "error above 5% of max|ref|"
            print(tag, "ok")
        except Exception as e:      # noqa: BLE001
            mxq.set_attention_path("tcgen05")
            bad += 1
            print(tag, "FAILED:", repr(e)[:300])
    # MX Linear
    for c in range(args.cases // 4):
        M, K, N = ri(1, 700), 64 * ri(1, 24), 4 * ri(1, 300)
        bfloat = 16 if ri(0, 1) else 32
        x = torch.randn(M, K, generator=rng) * torch.exp(0.5 * torch.randn(M, 1, generator=rng))
        w = torch.randn(N, K, generator=rng) * K ** -0.5
        b = torch.randn(N, generator=rng) * 0.1 if ri(0, 1) else None
        tag = f"linear {c}: M{M} K{K} N{N} bf{bfloat} bias{int(b is not None)}"
        try:
            y = mxq.mx_linear(x.cuda(), w.cuda(), None if b is None else b.cuda(), mx_specs(bfloat, False)).cpu()
            ref = O.mx_linear(x, w, b, bfloat=bfloat)
            err = (y - ref).abs()
            scale = float(ref.abs().max())
            if bfloat == 32:
                assert float(err.max()) <= 2e-5 * scale, float(err.max()) / scale
            else:
                assert float(err.max()) <= 2.0 ** -7 * scale and float((err > 0).float().mean()) <= 0.02
            print(tag, "ok")
        except Exception as e:      # noqa: BLE001
            bad += 1
            print(tag, "FAILED:", repr(e)[:300])
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
