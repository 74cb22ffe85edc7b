#!/bin/bash
# usage: bash tools/gpu_final.sh TAG   GPU tests + the default bench line + the ncu launch list (no full capture)
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log; tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python profiles/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | head -12 | cut -c1-150
bash tools/gpu_list.sh $TAG | head -12
