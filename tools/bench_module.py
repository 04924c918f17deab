#!/usr/bin/env python
"""A whole attention module of the workloads on B200: MX Linear qkv projection -> pruned MXINT8 attention
-> MX Linear output projection (the reference's QuantizedAttention.forward, workloads/deit/scripts/
main.py:85-157, with its mx.Linear projections), through the module shims.
    python tools/bench_module.py [--reps 20]
Each line also carries the module END TO END from host buffers (`e2e_*`): x pinned on the host -> device, the module, y back
to a pinned host buffer, double-buffered on three streams with the copies inside the timed region - one third of the bytes
the attention-only e2e of bench.py moves (x and y instead of q, k, v and out), which is what INTEGRATION.md integrates."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from bench import mx_specs  # noqa: E402
from mx_quantization_b200.modules import Attention, QuantizedAttention  # noqa: E402


class _TimmAttention(nn.Module):            # the attributes QuantizedAttention takes from timm's Attention
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.proj_drop = nn.Identity()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cases = [
        ("deit_base_attention", QuantizedAttention(_TimmAttention(768, 12), mx_quant=True, mx_specs=mx_specs(32, False),
                                                   top_k=True, k=30, approx_flag=True, pred_mode="ex_pred"), 256, 197, 768, 12),
        ("dit_xl2_attention", Attention(1152, num_heads=16, qkv_bias=True, mx_quant=True, mx_specs=mx_specs(16, False),
                                        top_k=True, k=154, ex_pred=True, pred_mode="ex_pred"), 256, 256, 1152, 16),
    ]
    for name, mod, B, N, C, H in cases:
        mod = mod.to(dev).eval()
        x = torch.randn(B, N, C, device=dev)
        with torch.no_grad():
            ms = timed(lambda: mod(x), args.reps)
        # end to end from host buffers: NB independent inputs, H2D / compute / D2H overlapped across iterations
        NB = 3
        xh = [torch.randn(B, N, C).pin_memory() for _ in range(NB)]
        yh = [torch.empty(B, N, C).pin_memory() for _ in range(NB)]
        xd = [torch.empty(B, N, C, device=dev) for _ in range(NB)]
        s_in, s_out, s_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(NB)]
        ev_c = [torch.cuda.Event() for _ in range(NB)]
        ev_out = [torch.cuda.Event() for _ in range(NB)]

        def e2e_iter(i):
            b = i % NB
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_c[b])            # the previous user of xd[b] has computed
                xd[b].copy_(xh[b], non_blocking=True)
                ev_in[b].record(s_in)
            s_c.wait_event(ev_in[b])
            with torch.no_grad():
                y = mod(xd[b])
            ev_c[b].record(s_c)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c[b])
                yh[b].copy_(y, non_blocking=True)
                y.record_stream(s_out)
                ev_out[b].record(s_out)

        for i in range(NB):
            e2e_iter(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.reps):
            e2e_iter(i)
        s_c.wait_stream(s_in)
        s_c.wait_stream(s_out)
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / args.reps
        nbytes = B * N * C * 4
        print(json.dumps({"module": name, "batch": B, "tokens": N, "dim": C, "heads": H, "ms": ms,
                          "tokens_per_s": B * N / (ms * 1e-3), "heads_per_s": B * H / (ms * 1e-3),
                          "e2e_ms": ems, "e2e_tokens_per_s": B * N / (ems * 1e-3), "e2e_heads_per_s": B * H / (ems * 1e-3),
                          "e2e_h2d_bytes": nbytes, "e2e_d2h_bytes": nbytes,
                          "e2e_h2d_gbs": nbytes / (ems * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
