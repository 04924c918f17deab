#!/bin/bash
# Long-sequence (C5) evidence on one B200: full GPU test suite, smoke, C5 sweep with the oracle check at every N, A/B of the
# selection kernel on 9 input kinds, ncu --set full of the two long-sequence kernels at N = 4096, a short default bench line.
#   gpurun --timeout 1200 -- 'bash tools/gpu_long.sh r02b'
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.txt
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.txt 2>&1; tail -1 gpurun_out/smoke_$TAG.txt
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/bench_${TAG}_quick.json 2> gpurun_out/bench_${TAG}_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_quick.json')); print('bench', round(d['ms_per_step'],3), 'ms/step', int(d['value']), d['unit'], d['clocks'])" || tail -3 gpurun_out/bench_${TAG}_quick.err
timeout 500 python tools/sweep_c5.py --reps 3 --check > gpurun_out/c5_sweep_$TAG.jsonl 2> gpurun_out/c5_sweep_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/c5_sweep_$TAG.jsonl"):
    d = json.loads(l)
    print(d["N"], d["ratio"], round(d["ms"], 3), "ms", int(d["heads_per_s"]), "heads/s", {k: round(v, 3) for k, v in d["kernel_ms"].items()},
          d.get("mask_bit_exact"), d.get("out_max_abs_err_rel"))
PY
timeout 400 python tools/ab_long_select.py --ns 512,1024,2048,4096 --reps 3 > gpurun_out/ab_long_$TAG.jsonl 2>&1; tail -1 gpurun_out/ab_long_$TAG.jsonl
timeout 60 python tools/prof_long.py 4096 0.25 2 > gpurun_out/plain_long_$TAG.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_select_long_tc|k_attend_long_pair" -s 2 -c 2 \
    -o gpurun_out/prof_long_$TAG python tools/prof_long.py 4096 0.25 2 > gpurun_out/ncu_long_$TAG.log 2>&1
tail -1 gpurun_out/plain_long_$TAG.log; tail -1 gpurun_out/ncu_long_$TAG.log
