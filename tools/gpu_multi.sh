#!/bin/bash
# One multi-GPU gpurun call (N = 2 / 4 / 8):  gpurun --gpus N --timeout T -- 'bash tools/gpu_multi.sh N TAG'
# two-device test, weak + strong scaling bench lines, C5 points, bare H2D/D2H copy ceiling.  Every step under `timeout`.
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 120 python -m pytest tests -m gpu -x -q -k "two_devices" 2>&1 | tail -3 > gpurun_out/pytest_2dev_${TAG}_${N}gpu.txt; cat gpurun_out/pytest_2dev_${TAG}_${N}gpu.txt
timeout 200 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/bench_${TAG}_weak_${N}gpu.json 2> gpurun_out/bench_${TAG}_weak_${N}gpu.err
timeout 200 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --no-others > gpurun_out/bench_${TAG}_strong_${N}gpu.json 2> gpurun_out/bench_${TAG}_strong_${N}gpu.err
timeout 200 $TR tools/sweep_c5.py --ns 256,1024,4096 --ratios 0.1,0.5 --reps 2 > gpurun_out/c5_${TAG}_${N}gpu.jsonl 2> gpurun_out/c5_${TAG}_${N}gpu.err
timeout 100 $TR tools/h2d_scaling.py > gpurun_out/h2d_${TAG}_${N}gpu.json 2> gpurun_out/h2d_${TAG}_${N}gpu.err
python - <<PY
import json
for kind in ("weak", "strong"):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_%s_${N}gpu.json" % kind).read().strip().splitlines()[-1])
        print(kind, d["n_gpus"], "gpus", int(d["value"]), d["unit"], round(d["ms_per_step"], 3), "ms/step e2e", d.get("e2e") and int(d["e2e"]["value"]), d.get("verification"))
    except Exception as e:
        print(kind, "failed", e); print(open("gpurun_out/bench_${TAG}_%s_${N}gpu.err" % kind).read()[-800:])
for f in ("gpurun_out/c5_${TAG}_${N}gpu.jsonl", "gpurun_out/h2d_${TAG}_${N}gpu.json"):
    try:
        for l in open(f):
            if not l.startswith("{"):
                continue
            d = json.loads(l)
            print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k not in ("ranks", "kernel_ms")})
    except Exception as e:
        print(f, "failed", e)
PY
