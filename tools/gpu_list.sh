#!/bin/bash
# usage: bash tools/gpu_list.sh TAG   ncu launch list (durations only) of two single-context keyframes -> per-kernel summary
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --quick --inflight 1 > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
python profiles/summarize_launches.py $OUT/launches_$TAG.csv 2>/dev/null | head -45
