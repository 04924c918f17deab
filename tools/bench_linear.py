#!/usr/bin/env python
"""MX Linear throughput on the qkv / proj / MLP shapes of the workloads, beside a plain
cuBLAS bf16 GEMM of the same shape (torch.matmul) as the tensor-core reference point.
    python tools/bench_linear.py [--reps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from bench import mx_specs  # noqa: E402

SHAPES = [  # name, tokens, in, out
    ("deit_base_qkv", 256 * 197, 768, 2304),
    ("deit_base_proj", 256 * 197, 768, 768),
    ("deit_base_fc1", 256 * 197, 768, 3072),
    ("deit_base_fc2", 256 * 197, 3072, 768),
    ("dit_xl2_qkv", 256 * 256, 1152, 3456),
]


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
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    specs = mx_specs(32, False)
    for name, M, K, N in SHAPES:
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * K ** -0.5
        b = torch.randn(N, device=dev) * 0.1
        w_op = mxq.mx_linear_prepare_weight(w, specs)
        ms = timed(lambda: mxq.mx_linear(x, w_op, b, specs, out_features=N), args.reps)
        xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
        ms_cublas = timed(lambda: torch.matmul(xb, wb.t()), args.reps)
        flops = 2.0 * M * N * K
        print(json.dumps({"shape": name, "M": M, "K": K, "N": N, "ms": ms, "tflops": flops / ms / 1e9,
                          "frac_of_bf16_peak": flops / ms / 1e9 / peak_tf, "cublas_bf16_gemm_ms": ms_cublas,
                          "cublas_tflops": flops / ms_cublas / 1e9,
                          "note": "ms includes the activation quantizer (fp32 -> MXINT8 -> bf16 operand) and the fp32 output write"}),
              flush=True)


if __name__ == "__main__":
    main()
