import sys, os, torch
sys.path.insert(0, "/root/repo")
import bench, mx_quantization_b200 as mxq
dev = torch.device("cuda:0")
specs = bench.mx_specs(32, False)
for (B, H, k) in ((8, 3, 80), (8, 3, 30), (4, 12, 30), (8, 12, 30), (12, 12, 30), (16, 12, 30), (24, 12, 30)):
    N, hd = 197, 64
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B, N, 3, H, hd, device=dev, generator=g).permute(2, 0, 3, 1, 4)
    out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
    res = {}
    for mode in (2, 0, 1):
        mxq.set_fused_path(mode)
        for _ in range(5): mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, k, out=out)
        n = mxq.last_launch_count()
        torch.cuda.synchronize()
        # replay from a graph to take the host out of the picture
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(10): mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, k, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): gr.replay()
        e0.record()
        for _ in range(10): gr.replay()
        e1.record(); torch.cuda.synchronize()
        res[mode] = (round(e0.elapsed_time(e1) * 10, 2), n)   # us per call
    mxq.set_fused_path(True)
    print(f"heads {B*H:4d} k {k}: us per call (launches) forced-fused {res[2]}, three {res[0]}, default {res[1]}", flush=True)
