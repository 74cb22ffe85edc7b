#!/bin/bash
# quick GPU iteration: parity tests, then short bench runs for a sweep of an environment knob
# usage: bash tools/gpu_quick.sh TAG [ENVVAR v1 v2 ...]
TAG=$1; shift
VAR=$1; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log
if [ -z "$VAR" ]; then
  python bench.py --steps 12 --warmup 3 --quick > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
  python profiles/show_bench.py $OUT/bench_$TAG.json | head -24
else
  for v in "$@"; do
    echo "=== $VAR=$v"
    env $VAR=$v python bench.py --steps 12 --warmup 3 --quick > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err; echo "bench rc=$?"
    python profiles/show_bench.py $OUT/bench_${TAG}_$v.json | head -12
  done
fi
