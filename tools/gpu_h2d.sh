#!/bin/bash
# bare copy ceiling only:  gpurun --gpus N --timeout 300 -- 'bash tools/gpu_h2d.sh N TAG'
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_scaling.py 2> gpurun_out/h2d_${TAG}_${N}gpu.err | grep "^{" > gpurun_out/h2d_${TAG}_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/h2d_${TAG}_${N}gpu.json')); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k!='ranks'})"
