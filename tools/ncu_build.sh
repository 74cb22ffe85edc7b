#!/bin/bash
# ncu --set full capture of the once-per-keyframe kernels (features, forest, lattice build).  usage: bash tools/ncu_build.sh TAG
TAG=${1:-build}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'tile_csr_build|lattice_embed|forest_frame|gradient_mask|dist_forward|dist_backward|cloud_kernel|lab_border|neighbors_kernel|remap_offsets|splat_ones|upsample|integral_wavefront|slice_kernel|assign_ids|first_bitmap|blur_coop_kernel|norm_kernel|feat_' \
    --launch-skip 90 --launch-count 32 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --quick --inflight 1 --repeats 1 > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
for k in tile_csr_build forest_frame lattice_embed; do
  ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$k > $OUT/prof_${TAG}_src_$k.csv 2>/dev/null
done
ls -la $OUT/prof_$TAG.ncu-rep; rm -f $OUT/prof_$TAG.ncu-rep; du -sh $OUT
