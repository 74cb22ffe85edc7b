#!/bin/bash
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
bash tools/gpu_list.sh $TAG | grep -i "integral\|launches,"
timeout 200 python bench.py --steps 16 --warmup 3 --quick --repeats 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python profiles/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | sed -n 2,8p | cut -c1-150
