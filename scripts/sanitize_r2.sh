#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the shared-atomic encoders, the mbarrier pipelines of the network kernels and the
# inflate kernel, at small sizes.  Run on a GPU box; logs land in gpurun_out/ and are copied to profiles/.
set -u
CS=/usr/local/cuda/bin/compute-sanitizer
OUT=gpurun_out
mkdir -p $OUT
run() {  # name tool pytest-args...
    local name=$1 tool=$2; shift 2
    timeout 900 $CS --tool $tool --print-limit 5 --error-exitcode 86 --log-file $OUT/r2_sanitizer_${name}_${tool}.log \
        python -m pytest -x -q "$@" > $OUT/r2_sanitizer_${name}_${tool}.pytest 2>&1
    echo "$name $tool exit=$? $(tail -1 $OUT/r2_sanitizer_${name}_${tool}.pytest)" | tee -a $OUT/r2_sanitizer_summary.txt
    tail -3 $OUT/r2_sanitizer_${name}_${tool}.log >> $OUT/r2_sanitizer_summary.txt
}
: > $OUT/r2_sanitizer_summary.txt
for tool in memcheck racecheck; do
    run variant_encoder $tool tests/test_variant_encoder_gpu.py -k "kats or empty_and_small"          # k_tile_count, k_collect_ops, k_site_alleles, k_windows
    run polish_encoder $tool tests/test_polish_encoder_gpu.py -k "kat"                                 # k_polish_count
    run nets_variant $tool "tests/test_nets_gpu.py::test_variant_net_vs_oracle[130-2-1]"                 # k_lstm_layer, k_tc_gemm_p
    run nets_polish $tool "tests/test_nets_gpu.py::test_polish_net_vs_oracle[5-5-1]"                     # k_gru_layer
    run inflate $tool tests/test_inflate_gpu.py -k "inflate_matches or rejects"                         # k_bgzf_inflate
done
cat $OUT/r2_sanitizer_summary.txt
