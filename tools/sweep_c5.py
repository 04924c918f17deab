#!/usr/bin/env python
"""BASELINE.json config 5: sequence-length sweep 256-4096 tokens (DiT-style, 16 heads, head_dim 72,
constant 65536 tokens per batch), top-k ratio 0.1-0.5.  Prints one JSON line per point:
heads/s, tokens/s, per-kernel ms, GB/s against the HBM roofline, and (--check) a bit-exact
mask check + output check of one head against the CPU oracle at EVERY N.
    python tools/sweep_c5.py [--reps 3] [--check]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_c5.py ...
        G GPUs: the constant-token batch (65536 / N entries) is cut into contiguous batch slices
        (mx_quantization_b200.sharding), no data-path collective; time = max over ranks, heads/s = all ranks' heads."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mx_quantization_b200 as mxq  # noqa: E402
from bench import bytes_per_head, hbm_peak, mx_specs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--ns", default="256,512,1024,2048,4096")
    ap.add_argument("--ratios", default="0.1,0.25,0.5")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        from mx_quantization_b200 import sharding
        dist.init_process_group("nccl", device_id=dev)
    specs = mx_specs(32, False)
    H, hd = 16, 72
    peak, _ = hbm_peak()
    for N in [int(x) for x in args.ns.split(",")]:
        Bg = max(1, 65536 // N)                     # global batch (constant tokens)
        g = torch.Generator(device=dev).manual_seed(N)
        buf = torch.randn(Bg, N, 3, H, hd, device=dev, generator=g)
        if world > 1:
            kind, lo, hi = sharding.shard_batch_heads(Bg, H, world, rank)
            assert kind == "batch", "sweep_c5: more ranks than batch entries"
            buf = buf[lo:hi].clone()
        B = buf.shape[0]
        qkv = buf.permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        out = torch.empty(B, N, H, hd, device=dev).permute(0, 2, 1, 3)
        for r in [float(x) for x in args.ratios.split(",")]:
            top_k = max(1, int(-(-r * N // 1)))
            for _ in range(2):
                mxq.pruned_attention(q, k, v, specs, top_k, out=out)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                mxq.pruned_attention(q, k, v, specs, top_k, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            km = []
            mxq.pruned_attention(q, k, v, specs, top_k, out=out, _kernel_ms=km)
            bph = bytes_per_head(N, hd)
            line = {"N": N, "B": Bg, "n_gpus": world, "B_per_gpu": B, "H": H, "hd": hd, "ratio": r, "top_k": top_k, "ms": ms,
                    "heads_per_s": Bg * H / (ms * 1e-3), "tokens_per_s": Bg * N / (ms * 1e-3),
                    "kernel_ms": {"predict_topk": km[0], "prep_v": km[1], "exact_attention": km[2]},
                    "predict_topk_gbs": bph["predict_topk"] * B * H / (km[0] * 1e-3) / 1e9,
                    "predict_topk_frac_of_hbm": bph["predict_topk"] * B * H / (km[0] * 1e-3) / 1e9 / peak,
                    "full_path_gbs": bph["full"] * B * H / (ms * 1e-3) / 1e9,
                    # at long N the predictor is bound by its N^2 nb block scorings, not by bytes - report that
                    # fraction too: (block scorings / s) against the 16 POPC lanes/clk/SM x SMs x clock the survey assumes
                    "predict_block_scorings_per_s": float(N) * N * ((hd + 31) // 32) * B * H / (km[0] * 1e-3),
                    "predict_frac_of_popc_issue_peak": float(N) * N * ((hd + 31) // 32) * B * H / (km[0] * 1e-3) / (16 * 148 * 1.965e9)}
            if args.check and rank == 0:
                from oracle import mxint8_oracle as O
                print("check: gpu slice", file=sys.stderr, flush=True)
                o2, mask = mxq.pruned_attention(q[:1, :1], k[:1, :1], v[:1, :1], specs, top_k, return_mask=True)
                torch.cuda.synchronize()
                print("check: oracle", file=sys.stderr, flush=True)
                ref = O.pruned_attention(q[:1, :1].cpu(), k[:1, :1].cpu(), v[:1, :1].cpu(), top_k, integer_scores=True)
                print("check: compare", file=sys.stderr, flush=True)
                want = O.mask_words_to_dense(O.idx_to_mask_words(ref["idx"], N), N)
                got = O.mask_words_to_dense(mask.cpu().to(torch.int64) & 0xFFFFFFFF, N)
                line["mask_bit_exact"] = bool(torch.equal(got, want))
                line["out_max_abs_err_rel"] = float((o2.cpu() - ref["out"]).abs().max() / ref["out"].abs().max())
            if rank == 0:
                print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
