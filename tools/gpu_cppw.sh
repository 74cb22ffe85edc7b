#!/bin/bash
# usage (gpurun --gpus 2): bash tools/gpu_cppw.sh TAG   validates the bench's cpp_worker leg at N = 2 (torchrun) and N = 1 (full default run)
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --quick > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err; echo "2 GPUs rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_2gpu_$TAG.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d.get("cpp_worker"))
PY
CUDA_VISIBLE_DEVICES=0 timeout 500 python bench.py --no-15m > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err; echo "1 GPU rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_1gpu_$TAG.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d.get("cpp_worker"))
PY
