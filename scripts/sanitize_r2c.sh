#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the reworked variant-encoder kernels (k_cigar_prefix, k_tile_count with difference
# arrays + block scan, flat k_collect_ops, shared-memory k_windows) at small sizes; logs land in gpurun_out/ (copied to profiles/).
set -u
CS=/usr/local/cuda/bin/compute-sanitizer
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/r2c_sanitizer_summary.txt
for tool in memcheck racecheck; do
    timeout 300 $CS --tool $tool --print-limit 5 --error-exitcode 86 --log-file $OUT/r2c_sanitizer_variant_encoder_${tool}.log \
        python -m pytest -x -q tests/test_variant_encoder_gpu.py -k "kats or empty_and_small" > $OUT/r2c_sanitizer_variant_encoder_${tool}.pytest 2>&1
    echo "variant_encoder $tool exit=$? $(tail -1 $OUT/r2c_sanitizer_variant_encoder_${tool}.pytest)" | tee -a $OUT/r2c_sanitizer_summary.txt
    tail -3 $OUT/r2c_sanitizer_variant_encoder_${tool}.log >> $OUT/r2c_sanitizer_summary.txt
done
cat $OUT/r2c_sanitizer_summary.txt
