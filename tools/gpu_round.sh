#!/bin/bash
# One gpurun call for a round's evidence: GPU parity tests, the bench line, the reference arm, the ncu launch list and a
# `--set full` capture of the hot kernels (exported to CSV on the box: gpurun merges at most 64 MiB back).
# Usage (repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG [noncu]'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$SKIPTEST" ]; then  # SKIPTEST=1: the same build was just tested by a separate call
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 30 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python profiles/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | head -30
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm rc=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --quick --inflight 1 > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'meanfield_point_kernel|blur_multi_coop|integral_wavefront|lattice_embed|tile_csr_build|forest_frame_lowres|upsample_|gradient_mask|neighbors_kernel' \
    --launch-skip 40 --launch-count 20 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --quick --inflight 1 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:meanfield_point_kernel > $OUT/prof_${TAG}_src_point.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:blur_multi_coop > $OUT/prof_${TAG}_src_blur.csv 2>/dev/null
ls -la $OUT/prof_$TAG.ncu-rep; [ $(stat -c %s $OUT/prof_$TAG.ncu-rep) -gt 40000000 ] && rm -f $OUT/prof_$TAG.ncu-rep
fi
du -sh $OUT
