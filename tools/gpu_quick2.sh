#!/bin/bash
# Quick A/B on one GPU: a parity subset + device-timed bench lines of the three workload shapes.
TAG=${1:-q}
KEXPR=${2:-"topk_mask or fused or end_to_end or cost_follows or golden or tight_lane or full_size"}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -4
for W in deit_base_c2 dit_xl2_c3 pixart_c4; do
  timeout 120 python bench.py --workload $W --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$W.json')); r=d['roofline']; print('$W', round(d['ms_per_step'],4), r['kernel'], r.get('avg_ms'), {k: round(v['avg_ms'],4) for k,v in r.get('three_kernel_path',r)['kernels'].items()})" || tail -5 gpurun_out/bench_${TAG}_$W.err
done
