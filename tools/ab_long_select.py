#!/usr/bin/env python
"""A/B of the long-sequence selection kernel (k_select_long_tc, Nk > 256): the sampled fine window in
front of the radix levels (default) against the radix levels alone (mxp_set_fused_path(0)).
Masks of the two must be IDENTICAL on every input kind (the radix levels are the ones pinned to the
oracle by tests/test_gpu_parity.py::test_long_sequence_end_to_end); prints one JSON line per point
with both timings of predict_topk (CUDA events, whole call = operand pre-pass + selection + the
CUDA-core pass over flagged rows).
    python tools/ab_long_select.py [--ns 512,1024,2048,4096] [--reps 3] [--check-only]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from bench import mx_specs  # noqa: E402

KINDS = ("randn", "lognormal0.25", "lognormal0.5", "lognormal1.5", "ties", "constant", "zeros", "skewed", "outliers")


def make(kind, B, H, N, hd, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.randn(B, H, N, hd, device=dev, generator=g)
    k = torch.randn(B, H, N, hd, device=dev, generator=g)
    if kind.startswith("lognormal"):
        s = float(kind[len("lognormal"):])
        q = q * torch.exp(s * torch.randn(B, H, N, 1, device=dev, generator=g))
        k = k * torch.exp(s * torch.randn(B, H, N, 1, device=dev, generator=g))
    elif kind == "ties":                    # few distinct key rows: long runs of equal scores
        k = k[:, :, :7].repeat(1, 1, (N + 6) // 7, 1)[:, :, :N].contiguous()
    elif kind == "constant":                # every key row the same: all scores of a row tie
        k = k[:, :, :1].expand(B, H, N, hd).contiguous()
    elif kind == "zeros":                   # zero keys in the second half, zero query rows here and there
        k[:, :, N // 2:] = 0.0
        q[:, :, ::17] = 0.0
    elif kind == "skewed":                  # the sample (first 256 keys) is not representative of the rest
        k[:, :, :256] *= 0.05
        k[:, :, 256:] += 0.5
    elif kind == "outliers":                # a handful of huge keys widen the static key window
        k[:, :, 5::511] *= 40.0
        q[:, :, 3::97, :32] *= 0.01
    return q, k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="512,1024,2048,4096")
    ap.add_argument("--ratios", default="0.1,0.25,0.5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check-only", action="store_true")
    ap.add_argument("--timing-heads", type=int, default=256, help="heads * N / 4096 of the timed randn call")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    specs = mx_specs(32, False)
    bad = 0
    for N in [int(x) for x in args.ns.split(",")]:
        for kind in KINDS:
            for hd in (72, 64):
                q, k = make(kind, 1, 3, N, hd, dev, seed=N + hd)
                for r in [float(x) for x in args.ratios.split(",")]:
                    top_k = max(1, int(r * N))
                    mxq.set_fused_path(0)
                    want = mxq.predict_topk(q, k, specs, top_k, return_idx=True)
                    mxq.set_fused_path(1)
                    got = mxq.predict_topk(q, k, specs, top_k, return_idx=True)
                    ok = torch.equal(want["mask"], got["mask"]) and torch.equal(want["idx"], got["idx"])
                    if r == 0.25:                       # and both against the CUDA-core integer kernel (no tensor core, no 15-bit window)
                        mxq.set_predict_path("cuda_core")
                        core = mxq.predict_topk(q, k, specs, top_k, return_idx=True)
                        mxq.set_predict_path("tcgen05")
                        ok = ok and torch.equal(core["mask"], got["mask"]) and torch.equal(core["idx"], got["idx"])
                    bad += 0 if ok else 1
                    if not ok:
                        rows = (want["mask"] != got["mask"]).any(-1).sum().item()
                        print(json.dumps({"N": N, "kind": kind, "hd": hd, "top_k": top_k, "equal": False, "rows_differ": rows}), flush=True)
        print(json.dumps({"N": N, "kinds": len(KINDS), "masks_and_idx_equal": bad == 0}), flush=True)
        if args.check_only:
            continue
        # timing on the C5 shape: 16 heads, head_dim 72, constant 65536 tokens per batch
        B = max(1, 65536 // N) * args.timing_heads // 256
        for kind in ("randn", "lognormal0.5"):
            q, k = make(kind, max(1, B), 16, N, 72, dev, seed=1)
            for r in [float(x) for x in args.ratios.split(",")]:
                top_k = max(1, int(r * N))
                ms = {}
                for mode in (0, 1):
                    mxq.set_fused_path(mode)
                    for _ in range(2):
                        mxq.predict_topk(q, k, specs, top_k)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.reps):
                        mxq.predict_topk(q, k, specs, top_k)
                    e1.record()
                    torch.cuda.synchronize()
                    ms[mode] = e0.elapsed_time(e1) / args.reps
                mxq.set_fused_path(1)
                print(json.dumps({"N": N, "kind": kind, "heads": q.shape[0] * 16, "top_k": top_k,
                                  "radix_levels_ms": round(ms[0], 4), "fine_window_ms": round(ms[1], 4),
                                  "speedup": round(ms[0] / ms[1], 3)}), flush=True)
    print("FAILURES", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
