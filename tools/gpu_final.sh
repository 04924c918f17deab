#!/bin/bash
# Round-end style run on one B200: smoke, default bench (+ reference arm), other workloads, ncu evidence.
TAG=${1:-r01}
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.txt 2>&1; tail -2 gpurun_out/smoke_$TAG.txt
( time python bench.py ) > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; tail -4 gpurun_out/bench_${TAG}_default.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -4 gpurun_out/bench_${TAG}_reference.err
for W in dit_xl2_c3 pixart_c4 deit_tiny_c1; do
  python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d.get("roofline", {}).get("three_kernel_path", {}).get("kernels", {})
        print(f.split("/")[-1], d.get("impl", "ours"), int(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 3),
              "e2e", d.get("e2e") and int(d["e2e"]["value"]), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1),
              {k: (round(v["avg_ms"], 3), round(v["frac"], 3)) for k, v in ks.items()}, d.get("clocks"))
    except Exception as e:
        print(f, "failed:", e)
PY
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_launches_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/prof_layer.py deit_base_c2 3 > gpurun_out/plain_full_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_predict_topk_tc|k_attend_pair|k_attend_umma|k_prep_v" -s 3 -c 3 \
    -o gpurun_out/prof_full_$TAG python tools/prof_layer.py deit_base_c2 3 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
