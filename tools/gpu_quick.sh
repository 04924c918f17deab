#!/bin/bash
# Quick A/B on one GPU: fused-path parity tests + the DeiT-base bench line (device-timed only).
#   gpurun --timeout 420 -- 'bash tools/gpu_quick.sh TAG [pytest -k expr]'
TAG=${1:-q}
KEXPR=${2:-"fused or end_to_end or cost_follows or golden or tight_lane"}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -4
for W in deit_base_c2; do
  timeout 120 python bench.py --workload $W --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$W.json')); r=d['roofline']; print('$W', round(d['ms_per_step'],4), r['kernel'], r.get('avg_ms'), {k: round(v['avg_ms'],4) for k,v in r['three_kernel_path']['kernels'].items()})" || tail -5 gpurun_out/bench_${TAG}_$W.err
done
