#!/bin/bash
# One gpurun call: GPU parity tests, then the bench on both headline workloads (no CPU leg).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG [pytest -k expr]'
TAG=${1:-dev}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt
else
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt
fi
tail -12 gpurun_out/pytest_gpu_$TAG.txt
for W in deit_base_c2 dit_xl2_c3; do
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others \
      > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$W.json"))
    r = d["roofline"]
    ks = r["three_kernel_path"]["kernels"]
    print("$W", int(d["value"]), "heads/s", round(d["ms_per_step"], 3), "ms/step |", r["kernel"], round(r["avg_ms"], 4), "ms frac", round(r["frac"], 3),
          "| full path frac", round(r["full_path_frac"], 3), "| 3-kernel", {k: (round(v["avg_ms"], 3), round(v["frac"], 3)) for k, v in ks.items()})
except Exception as e:
    print("$W bench failed:", e)
    print(open("gpurun_out/bench_${TAG}_$W.err").read()[-1500:])
PY
done
