#!/bin/bash
# ncu --set full capture of the mean-field kernels only (warm caches: --cache-control none).  usage: bash tools/ncu_mf.sh TAG [regex]
TAG=${1:-mf}
RE=${2:-meanfield_point_kernel|blur_multi_coop}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$RE" \
    --launch-skip 44 --launch-count 6 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --quick --inflight 1 --repeats 1 > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null; wc -c $OUT/prof_${TAG}_raw.csv
