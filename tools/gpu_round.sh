#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one `--set full` capture of the
# mean-field kernels.  Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python profiles/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | head -40
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --quick > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'splat_kernel|blur_coop_kernel|slice_softmax_kernel|integral_wavefront|lattice_embed|patch_features|csr_fill|remap_offsets|upsample_kernel|unary_from' \
    --launch-skip 200 --launch-count 40 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --quick > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
