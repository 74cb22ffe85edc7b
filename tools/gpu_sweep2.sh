#!/bin/bash
# usage: bash tools/gpu_sweep2.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...   (one short bench per environment setting)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
i=0
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg python bench.py --steps 12 --warmup 3 --quick > $OUT/bench_${TAG}_$i.json 2> $OUT/bench_${TAG}_$i.err; echo "bench rc=$?"
  python profiles/show_bench.py $OUT/bench_${TAG}_$i.json 2>/dev/null | sed -n 2,8p
  i=$((i+1))
done
