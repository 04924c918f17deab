#!/bin/bash
# Round-end style run on one B200: smoke, default bench (+ reference arm), other workloads, ncu evidence, C5 sweep,
# quantizer / module benches.  Every step under `timeout`.   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.txt 2>&1; tail -2 gpurun_out/smoke_$TAG.txt
( time timeout 500 python bench.py ) > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; tail -4 gpurun_out/bench_${TAG}_default.err
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -4 gpurun_out/bench_${TAG}_reference.err
for W in dit_xl2_c3 pixart_c4 deit_tiny_c1; do
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-others > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d.get("roofline", {}).get("three_kernel_path", {}).get("kernels", {})
        print(f.split("/")[-1], d.get("impl", "ours"), int(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 3),
              "e2e", d.get("e2e") and int(d["e2e"]["value"]), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1),
              d.get("roofline", {}).get("kernel"), d.get("roofline", {}).get("frac"),
              {k: (round(v["avg_ms"], 3), round(v["frac"], 3)) for k, v in ks.items()}, d.get("clocks"))
    except Exception as e:
        print(f, "failed:", e)
PY
timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others > gpurun_out/plain_launches_$TAG.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 120 python tools/prof_layer.py deit_base_c2 3 > gpurun_out/plain_full_$TAG.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused" -s 2 -c 1 \
    -o gpurun_out/prof_fused_$TAG python tools/prof_layer.py deit_base_c2 3 > gpurun_out/ncu_fused_$TAG.log 2>&1
timeout 120 python tools/prof_layer.py deit_base_c2 3 three > gpurun_out/plain_three_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_predict_topk_tc|k_attend_pair|k_prep_v" -s 3 -c 3 \
    -o gpurun_out/prof_three_$TAG python tools/prof_layer.py deit_base_c2 3 three > gpurun_out/ncu_three_$TAG.log 2>&1
tail -2 gpurun_out/ncu_three_$TAG.log
timeout 500 python tools/sweep_c5.py --reps 3 --check > gpurun_out/c5_sweep_$TAG.jsonl 2> gpurun_out/c5_sweep_$TAG.err; tail -2 gpurun_out/c5_sweep_$TAG.jsonl | cut -c1-400
timeout 200 python tools/bench_quant.py > gpurun_out/quant_$TAG.jsonl 2> gpurun_out/quant_$TAG.err; tail -2 gpurun_out/quant_$TAG.err
timeout 200 python tools/bench_module.py > gpurun_out/module_$TAG.jsonl 2> gpurun_out/module_$TAG.err; cat gpurun_out/module_$TAG.jsonl | cut -c1-300
