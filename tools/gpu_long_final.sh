#!/bin/bash
# Final long-sequence evidence of a build: full GPU suite, C5 sweep with the oracle check at every N, ncu --set full of the two
# long-sequence kernels at N = 4096.   gpurun --timeout 900 -- 'bash tools/gpu_long_final.sh r02d'
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.txt
timeout 500 python tools/sweep_c5.py --reps 3 --check > gpurun_out/c5_sweep_$TAG.jsonl 2> gpurun_out/c5_sweep_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/c5_sweep_$TAG.jsonl"):
    d = json.loads(l)
    print(d["N"], d["ratio"], round(d["ms"], 3), "ms", int(d["heads_per_s"]), "heads/s", {k: round(v, 3) for k, v in d["kernel_ms"].items()},
          d.get("mask_bit_exact"), d.get("out_max_abs_err_rel"), round(d["predict_topk_frac_of_hbm"], 4), round(d["predict_frac_of_popc_issue_peak"], 3))
PY
timeout 60 python tools/prof_long.py 4096 0.25 2 > gpurun_out/plain_long_$TAG.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_select_long_tc|k_attend_long_pair" -s 2 -c 2 \
    -o gpurun_out/prof_long_$TAG python tools/prof_long.py 4096 0.25 2 > gpurun_out/ncu_long_$TAG.log 2>&1
tail -1 gpurun_out/plain_long_$TAG.log; tail -1 gpurun_out/ncu_long_$TAG.log
