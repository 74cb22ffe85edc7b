#!/bin/bash
# ncu --set full of the mean-field kernels of one local map (sorted fused path).  usage: bash tools/ncu_map.sh TAG [points]
TAG=${1:-map}; NP=${2:-2000000}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'meanfield_point_kernel|blur_multi_coop|tile_csr_build|gather_rows' \
    --launch-skip 2 --launch-count 8 -o $OUT/prof_$TAG -f python tools/diag/local_map_time.py $NP > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:meanfield_point_kernel > $OUT/prof_${TAG}_src_point.csv 2>/dev/null
rm -f $OUT/prof_$TAG.ncu-rep; du -sh $OUT
