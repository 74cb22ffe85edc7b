#!/bin/bash
# One gpurun call for a round's evidence: GPU parity tests, the bench line, the ncu launch list and `--set full`
# captures of the dominant kernels.  Usage (repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python profiles/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | head -24
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm rc=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --quick --inflight 1 > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'meanfield_point_kernel|blur_multi_coop|integral_wavefront|lattice_embed|bitmap_prefix|tile_csr_build|forest_frame_lowres|upsample_kernel|remap_offsets|assign_ids' \
    --launch-skip 60 --launch-count 36 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --quick --inflight 1 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
