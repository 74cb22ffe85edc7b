#!/bin/bash
# usage: bash tools/gpu_sweep5.sh TAG "DEFS|bench args" ...   like gpu_sweep4.sh, but only the CRF tests on the first build
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0; last="@"
for spec in "$@"; do
  defs="${spec%%|*}"; extra=""; [[ "$spec" == *"|"* ]] && extra="${spec#*|}"
  echo "=== [$i] $defs | $extra"
  if [ "$defs" != "$last" ]; then RSS_NVCC_DEFS="$defs" python -m rovinasemanticsegmentation_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED; last="$defs"; fi
  if [ $i -eq 0 ]; then timeout 300 python -m pytest tests/test_gpu_crf.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -2; fi
  timeout 300 python bench.py --steps 16 --warmup 3 --quick --repeats 3 $extra > $OUT/bench_${TAG}_$i.json 2> $OUT/bench_${TAG}_$i.err; echo "bench rc=$?"
  python profiles/show_bench.py $OUT/bench_${TAG}_$i.json 2>/dev/null | sed -n 2,9p | cut -c1-130
  i=$((i+1))
done
