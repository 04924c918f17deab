#!/usr/bin/env python
"""bench.py - pruned-attention throughput of the MXINT8 exponent-sign hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (q,k,v -> out) over the workload's layers, each layer with
its own synthetic q/k/v already resident in HBM (independent tensors per layer, far larger than
L2, so no L2 flush is needed between iterations).  Prints ONE JSON line (rank 0).

Default workload = BASELINE.json configs[1]: DeiT-base attention, batch 256 x 197 tokens,
12 heads, head_dim 64, all 12 layers, k = 30 (workloads/deit/scripts/run_deit.sh:51).
Multi-GPU: batch x heads shards across ranks with no data-path collective (weak scaling: every
rank runs the full per-GPU workload); NCCL is used for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: B, H, N, hd, top_k, layers, bfloat, flush     (SURVEY.md section 8 config shapes)
    "deit_base_c2": dict(B=256, H=12, N=197, hd=64, top_k=30, layers=12, bfloat=32, flush=False),
    "dit_xl2_c3": dict(B=256, H=16, N=256, hd=72, top_k=154, layers=28, bfloat=16, flush=False),
    "pixart_c4": dict(B=256, H=16, N=256, hd=72, top_k=77, layers=28, bfloat=32, flush=True),
    "deit_tiny_c1": dict(B=8, H=3, N=197, hd=64, top_k=80, layers=1, bfloat=32, flush=False),
}
CPU_SAMPLE_B = 8          # cpu_baseline / reference arm: a B=8 slice of one layer


def mx_specs(bfloat, flush):
    return {
        'w_elem_format': 'int8', 'a_elem_format': 'int8', 'scale_bits': 8, 'shared_exp_method': 'max',
        'block_size': 32, 'bfloat': bfloat, 'fp': 0, 'bfloat_subnorms': True, 'round': 'nearest',
        'round_mx_output': 'nearest', 'round_output': 'nearest', 'round_weight': 'nearest',
        'mx_flush_fp32_subnorms': flush, 'custom_cuda': False, 'quantize_backprop': False,
    }


def bytes_per_head(N, hd):
    """ALGORITHMIC bytes per (batch, head) unit (DESIGN.md 'Measurement')."""
    nb, nw = (hd + 31) // 32, (N + 31) // 32
    hdp = (hd + 15) // 16 * 16
    pred = 2 * N * hd * 4 + N * nw * 4                       # Q,K fp32 in ; bitmask out
    prep = N * hd * 4 + N * hdp * 2                          # V fp32 in ; bf16 V^T operand out
    attn = 3 * N * hdp * 2 + N * nw * 4 + N * hd * 4         # bf16 Q,K,V operands + mask in ; O fp32 out
    full = 16 * N * hd                                       # Q,K,V in ; O out
    return {"predict_topk": pred, "prep_v": prep, "exact_attention": attn, "full": full}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_heads_per_s(w, budget_s=12.0, threads=None):
    """The oracle port of the reference's CPU path (torch.topk left in place, as the reference
    does), all host threads, on a B=CPU_SAMPLE_B slice of one layer."""
    import torch
    from oracle import mxint8_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = min(CPU_SAMPLE_B, w["B"])
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, w["H"], w["N"], w["hd"], generator=g) for _ in range(3))

    def run():
        O.pruned_attention(q, k, v, w["top_k"], bfloat=w["bfloat"], flush=w["flush"], use_torch_topk=True)

    run()                                           # warm-up
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter(); run(); times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return B * w["H"] / med, threads, f"B={B} slice of one layer ({B * w['H']} heads), median of {len(times)} reps"


def config_dict(name, w, gpus):
    return {"workload": name, "batch": w["B"], "heads": w["H"], "tokens": w["N"], "head_dim": w["hd"],
            "top_k": w["top_k"], "layers": w["layers"], "mx_specs": f"int8/block32/bfloat{w['bfloat']}"
            + ("/flush" if w["flush"] else ""), "parallelism": f"batch x heads sharded over {gpus} GPU(s), no collective",
            "l2": "inputs (one q/k/v set per layer) exceed L2; no flush needed"}


def run_reference(args, name, w):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the
    Python reference itself cannot travel to the GPU box), host cores only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import mxint8_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = min(CPU_SAMPLE_B, w["B"])
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, w["H"], w["N"], w["hd"], generator=g) for _ in range(3))

    def step():
        O.pruned_attention(q, k, v, w["top_k"], bfloat=w["bfloat"], flush=w["flush"], use_torch_topk=True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    heads = B * w["H"]
    val = heads / dt
    sample = f"each step = B={B} slice of one layer ({heads} heads) of {name}"
    print(json.dumps({
        "impl": "reference", "metric": "pruned-attention heads/s", "value": val, "unit": "heads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": config_dict(name, w, args.gpus), "tokens_per_s": B * w["N"] / dt,
        "cpu_baseline": {"value": val, "unit": "heads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "heads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="deit_base_c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    name, w = args.workload, WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, name, w)
        return

    import torch
    import torch.distributed as dist
    import mx_quantization_b200 as mxq

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    specs = mx_specs(w["bfloat"], w["flush"])
    B, H, N, hd, L, top_k = w["B"], w["H"], w["N"], w["hd"], w["layers"], w["top_k"]
    heads_per_step = B * H * L                      # per GPU (weak scaling)

    # synthetic activations: one fused qkv buffer per layer, q/k/v are the permuted views the
    # attention modules produce (workloads/deit/scripts/main.py:87-88)
    g = torch.Generator(device=dev).manual_seed(1000 * rank)
    layers = []
    for _ in range(L):
        buf = torch.randn(B, N, 3, H, hd, device=dev, generator=g)
        qkv = buf.permute(2, 0, 3, 1, 4)
        layers.append((qkv[0], qkv[1], qkv[2]))
    out = torch.empty(B, N, H, hd, device=dev)      # (B,N,H,hd): the module's transpose(1,2) is free
    out_view = out.permute(0, 2, 1, 3)

    def step():
        for (q, k, v) in layers:
            mxq.pruned_attention(q, k, v, specs, top_k, out=out_view)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    launches_per_call = mxq.last_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    # ---- per-kernel timing for the roofline: CUDA events around each kernel of the same call,
    # recorded inside the library on the launching stream (mxp_pruned_attention_profile)
    kt = {"predict_topk": 0.0, "prep_v": 0.0, "exact_attention": 0.0}
    reps = 0
    for it in range(1 + max(1, min(args.steps, 3))):        # first pass untimed
        for (q, k, v) in layers:
            ms3 = []
            mxq.pruned_attention(q, k, v, specs, top_k, out=out_view, _kernel_ms=ms3)
            if it == 0:
                continue
            kt["predict_topk"] += ms3[0]
            kt["prep_v"] += ms3[1]
            kt["exact_attention"] += ms3[2]
            reps += 1
    if rank == 0:
        # the timed region (a few tens of ms) can be shorter than one nvidia-smi period: keep the SAME
        # step loop running, untimed, until a few samples have been taken under that load
        t_end = time.perf_counter() + 3.0
        while len(sampler.lines) < 4 and time.perf_counter() < t_end:
            step()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = ("sampled from warm-up to the end of an untimed continuation of the same step loop "
                          "(nvidia-smi period 50 ms; the timed region alone can be shorter than one period)")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * heads_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timing
    e2e = None
    if not args.no_e2e:
        e2e_layers = min(L, 4)                      # bounded pinned footprint; per-layer cost is uniform
        host_in = [torch.empty(B, N, 3, H, hd).pin_memory() for _ in range(e2e_layers)]
        for hbuf in host_in:
            hbuf.normal_()
        host_out = [torch.empty(B, N, H, hd).pin_memory() for _ in range(e2e_layers)]
        NBUF = 3                                    # device staging buffers: keeps the H2D engine busy back to back
        dev_in = [torch.empty(B, N, 3, H, hd, device=dev) for _ in range(NBUF)]
        dev_out = [torch.empty(B, N, H, hd, device=dev) for _ in range(NBUF)]
        s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(NBUF)]
        ev_cmp = [torch.cuda.Event() for _ in range(NBUF)]
        ev_free = [torch.cuda.Event() for _ in range(NBUF)]
        ev_out = [torch.cuda.Event() for _ in range(NBUF)]

        def e2e_step():
            for li in range(e2e_layers):
                s = li % NBUF
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_free[s])                       # compute finished with dev_in[s]
                    dev_in[s].copy_(host_in[li], non_blocking=True)
                    ev_in[s].record(s_in)
                s_cmp.wait_event(ev_in[s])
                s_cmp.wait_event(ev_out[s])                           # previous D2H of dev_out[s] done
                qkv = dev_in[s].permute(2, 0, 3, 1, 4)
                mxq.pruned_attention(qkv[0], qkv[1], qkv[2], specs, top_k, out=dev_out[s].permute(0, 2, 1, 3))
                ev_cmp[s].record(s_cmp)
                ev_free[s].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[s])
                    host_out[li].copy_(dev_out[s], non_blocking=True)
                    ev_out[s].record(s_out)
            s_cmp.wait_stream(s_out)

        e2e_step()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        n_e2e = max(1, min(args.steps, 3))
        t0.record()
        for _ in range(n_e2e):
            e2e_step()
        t1.record()
        barrier()
        ems = t0.elapsed_time(t1) / n_e2e
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        scale_l = L / e2e_layers                    # bytes/heads reported for the full L-layer step
        e2e = {"value": world * B * H * e2e_layers / (ems * 1e-3), "unit": "heads/s",
               "h2d_bytes_per_step": int(host_in[0].numel() * 4 * e2e_layers * scale_l),
               "d2h_bytes_per_step": int(host_out[0].numel() * 4 * e2e_layers * scale_l),
               "ms_per_step": ems * scale_l,
               "note": f"timed on {e2e_layers} of {L} layers per step (uniform per-layer cost; pinned host "
                       "buffers, H2D/compute/D2H triple-buffered on 3 streams), scaled to the full step"}

    if rank == 0:
        peak, peak_src = hbm_peak()
        bph = bytes_per_head(N, hd)
        kernels = {}
        for kname, tot in kt.items():
            avg_ms = tot / reps
            if avg_ms <= 0:
                continue
            ach = bph[kname] * B * H / (avg_ms * 1e-3) / 1e9
            kernels[kname] = {"avg_ms": avg_ms, "achieved_gbs": ach, "frac": ach / peak,
                              "bytes_per_launch": bph[kname] * B * H}
        dom = max(kernels, key=lambda n: kernels[n]["avg_ms"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(name, {}).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "kernels": kernels,
                    "full_path_gbs": bph["full"] * heads_per_step / (ms_per_step * 1e-3) / 1e9}
        line = {
            "metric": "pruned-attention heads/s", "value": value, "unit": "heads/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config_dict(name, w, world),
            "tokens_per_s": world * B * N * L / (ms_per_step * 1e-3),
            "roofline": roofline, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches_per_call * L * args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_port_heads_per_s(w)
            line["cpu_baseline"] = {"value": v, "unit": "heads/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
