"""SURVEY.md 2a's bar for the quantizer alone: the reference's own CUDA kernels (quantize_mx_innermost_cuda_kernel for
head_dim 64, quantize_mx_by_tile_cuda_kernel for head_dim 72; microxscaling/mx/cpp/mx.cuh:57-158, built for sm_100a from
where they lie by oracle/ref_build/Makefile -> oracle/_ref/libmxref_cuda.so) timed beside mxp_quantize_mxint8 on the same
B200, on the Q tensors of the bench workloads.  Prints one JSON line per shape.

    python tools/bench_quant.py > profiles/r02_quantizer_vs_reference_cuda_b200.jsonl

The reference kernels write fake-quantised fp32 (4 B in + 4 B out per element); mxp_quantize_mxint8 writes int8 codes
+ int8 block exponents (+ sign words), i.e. the compact form the reference never materialises.  Values are compared:
dequantised codes == the reference kernel's output wherever the two exponent policies agree (the C++ path reads the
exponent BITS of the block maximum, the Python path - the golden one, which libmxprune follows - evaluates
floor(log2(.)) in fp32 and rounds up within a few ulps below a power of two)."""
import ctypes
import json
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402

ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmxref_cuda.so"))
ref.ref_cuda_quantize_innermost.argtypes = [c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]
ref.ref_cuda_quantize_by_tile.argtypes = [c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                          c_void_p, c_void_p]
dev = torch.device("cuda:0")
peak, _ = bench.hbm_peak()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name in ("deit_base_c2", "dit_xl2_c3"):
    w = bench.WORKLOADS[name]
    B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
    g = torch.Generator(device=dev).manual_seed(0)
    # three rotating inputs (each 155 / 302 MB > L2), contiguous (B,H,N,hd) as the reference kernels require
    xs = [torch.randn(B, H, N, hd, device=dev, generator=g) for _ in range(3)]
    out = torch.empty_like(xs[0])
    specs = bench.mx_specs(32, False)
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    total = xs[0].numel()
    it = [0]

    def run_ref():
        x = xs[it[0] % 3]; it[0] += 1
        if hd % 32 == 0:
            rc = ref.ref_cuda_quantize_innermost(c_void_p(x.data_ptr()), total, 32, 0, c_void_p(out.data_ptr()), st)
        else:
            rc = ref.ref_cuda_quantize_by_tile(c_void_p(x.data_ptr()), total // hd, hd, 1, 32, 0, c_void_p(out.data_ptr()), st)
        assert rc == 0, rc

    def run_ours():
        x = xs[it[0] % 3]; it[0] += 1
        return mxq.quantize_mxint8(x, specs)

    ms_ref, ms_ours = timed(run_ref), timed(run_ours)
    # value agreement on the last input the reference kernel processed
    it[0] = 0
    run_ref()
    codes, exps = mxq.quantize_mxint8(xs[0], specs)
    torch.cuda.synchronize()
    e = exps.to(torch.int32).repeat_interleave(32, dim=-1)[..., :hd]
    deq = torch.ldexp(codes.to(torch.float32), e - 6)
    same = float((deq == out).float().mean())
    nb = (hd + 31) // 32
    line = {"workload": name, "tensor": f"Q ({B},{H},{N},{hd}) fp32 contiguous", "elements": total,
            "reference_cuda_kernel": "quantize_mx_innermost_cuda_kernel" if hd % 32 == 0 else "quantize_mx_by_tile_cuda_kernel",
            "reference_ms": ms_ref, "reference_gbs": total * 8 / (ms_ref * 1e-3) / 1e9,
            "mxp_quantize_mxint8_ms": ms_ours, "mxp_gbs": (total * 5 + total // hd * nb) / (ms_ours * 1e-3) / 1e9,
            "speedup": ms_ref / ms_ours, "hbm_peak_gbs": peak,
            "reference_frac_of_peak": total * 8 / (ms_ref * 1e-3) / 1e9 / peak,
            "mxp_read_frac_of_peak": total * 4 / (ms_ours * 1e-3) / 1e9 / peak,
            "elements_equal_frac": same,
            "note": "reference: 4 B read + 4 B written per element (fake-quant fp32); mxp: 4 B read + 1 B code + exponents written; "
                    "elements differ only where the block maximum is within a few ulps below a power of two (exponent policy)"}
    print(json.dumps(line), flush=True)
