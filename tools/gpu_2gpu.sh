#!/bin/bash
# usage (gpurun --gpus 2): bash tools/gpu_2gpu.sh TAG   keyframes and local maps on 2 GPUs (one rank per GPU, torchrun), then 1 GPU local maps
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --quick > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err; echo "keyframes 2 GPUs rc=$?"
tail -1 $OUT/bench_2gpu_$TAG.json | cut -c1-260
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --workload local_maps --steps 8 --warmup 3 > $OUT/bench_lm_2gpu_$TAG.json 2> $OUT/bench_lm_2gpu_$TAG.err; echo "local maps 2 GPUs rc=$?"
tail -1 $OUT/bench_lm_2gpu_$TAG.json | cut -c1-260
timeout 300 python bench.py --workload local_maps --steps 8 --warmup 3 > $OUT/bench_lm_1gpu_$TAG.json 2> $OUT/bench_lm_1gpu_$TAG.err; echo "local maps 1 GPU rc=$?"
tail -1 $OUT/bench_lm_1gpu_$TAG.json | cut -c1-260
