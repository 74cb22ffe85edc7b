#!/bin/bash
# usage: bash tools/gpu_sweep_mf.sh TAG "DEFS1" "DEFS2" ...   rebuilds librss.so on the box with each -D set; short latency-focused bench each
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for defs in "$@"; do
  echo "=== $defs"
  RSS_NVCC_DEFS="$defs" python -m rovinasemanticsegmentation_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  if [ $i -eq 0 ]; then python -m pytest tests/test_gpu_crf.py -m gpu -x -q 2>&1 | tail -1; fi
  python bench.py --steps 12 --warmup 3 --quick --repeats 3 > $OUT/bench_${TAG}_$i.json 2> $OUT/bench_${TAG}_$i.err; echo "bench rc=$?"
  python profiles/show_bench.py $OUT/bench_${TAG}_$i.json 2>/dev/null | sed -n 2,6p | cut -c1-150
  i=$((i+1))
done
