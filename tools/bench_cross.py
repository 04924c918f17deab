#!/usr/bin/env python
"""PixArt-alpha cross-attention throughput: 256 latent tokens x 120 text tokens, 16 heads,
head_dim 72, additive text mask (1 - mask) * -10000 with ~30 valid tokens per sample, top-k 77
(workloads/PixArt/scripts/run_pixart_alpha.sh:27), batch 256.  Also the same shape without a mask.
    python tools/bench_cross.py [--reps 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from bench import mx_specs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, Nq, S, hd, top_k = 256, 16, 256, 120, 72, 77
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B, Nq, H, hd, device=dev, generator=g).permute(0, 2, 1, 3)
    kv = torch.randn(B, S, 2, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    valid = torch.randint(20, 41, (B,), device=dev, generator=g)
    mask = (torch.arange(S, device=dev)[None, :] < valid[:, None]).float()
    bias = ((1.0 - mask) * -10000.0).reshape(B, 1, 1, S)
    out = torch.empty(B, Nq, H, hd, device=dev).permute(0, 2, 1, 3)
    specs = mx_specs(32, True)
    for name, kb in (("masked", bias), ("no_mask", None)):
        def call():
            mxq.pruned_attention(q, k, v, specs, top_k, out=out, key_bias=kb)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(json.dumps({"workload": "pixart_cross_attention", "variant": name, "B": B, "H": H, "Nq": Nq, "Nk": S,
                          "hd": hd, "top_k": top_k, "ms": ms, "heads_per_s": B * H / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
