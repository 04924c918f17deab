"""Per-phase cycle accounting of the fused kernel (mxp_debug_fused_timing): where a group's thread 0 spends its time.
    make -C mx_quantization_b200/csrc timing && MXPRUNE_LIB=.../libmxprune_timing.so timeout 120 python tools/fused_timing.py [workload] [calls]
NOTE (round 2): the accounting build (-DMXP_FUSED_TIMING) DEADLOCKED on B200 in its last two runs (the command was killed by its
time limit) and was not debugged; it also covers only the run-time head_dim instantiations (mxprune_fused64.cu is linked without
the macro).  The per-phase shares under profiles/ come from tools/ncu_phases.py (ncu's per-instruction counts and stall samples)
instead.  Run this only under `timeout`."""
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mx_quantization_b200 as mxq  # noqa: E402
from mx_quantization_b200 import _lib  # noqa: E402

NAMES = {0: "p1 step: before TMA wait (issue next, loop)", 1: "p1 step: TMA wait", 2: "p1 step: quantize",
         3: "p1 step: barrier after quantize", 4: "p1 tile: MMA issue + row parameters", 5: "p1 tile: MMA wait",
         6: "p1 tile: keys TMEM->regs", 7: "p1 tile: barrier after keys", 8: "p1 tile: bisection", 9: "p1 tile: emit + generic rows",
         10: "p2 tile: TMA issue + mask loads", 11: "p2 tile: TMA wait", 12: "p2 tile: S MMA + wait", 13: "p2 tile: walk (compaction)",
         14: "p2 tile: pass B + window exps", 15: "p2 tile: barrier before pass C", 16: "p2 group: zero + scatter (+ wait prev MMA)",
         17: "p2 group: fence + barrier", 18: "p2 tile: MMA issue + final wait", 19: "p2 tile: O readout + stores",
         20: "p2 tile: end barrier", 21: "phase 1 -> 2 switch (fences + barrier)", 22: "tail"}

name = sys.argv[1] if len(sys.argv) > 1 else "deit_base_c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
B, H, N, hd = w["B"], w["H"], w["N"], w["hd"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
specs = bench.mx_specs(w["bfloat"], w["flush"])
for _ in range(2):
    mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out)
buf = torch.zeros(320 * 32, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.mxp_debug_fused_timing(c_void_p(buf.data_ptr()))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(calls):
    buf.zero_()
    mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, w["top_k"], out=out)
ev1.record()
torch.cuda.synchronize()
lib.mxp_debug_fused_timing(c_void_p(0))
t = buf.cpu().reshape(320, 32).double()
t = t[t.sum(1) > 0]
tot = t.sum(1).mean()
print(f"{name}: {ev0.elapsed_time(ev1) / calls:.3f} ms per call (with accounting), {t.shape[0]} groups, "
      f"{tot:.0f} cycles per group, {B * H / t.shape[0]:.2f} heads per group")
for i in range(23):
    m = t[:, i].mean()
    print(f"  [{i:2d}] {100 * m / tot:5.1f} %  {m / (B * H / t.shape[0]):9.0f} cyc/head   {NAMES[i]}")
