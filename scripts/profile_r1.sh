#!/bin/bash
# Run under gpurun (1 GPU).  Writes ncu captures into gpurun_out/; summaries are copied to profiles/ afterwards.
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
# every launch with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/launches_r1.log 2>&1
# the pileup-count kernel, full set
ncu --set full --clock-control none --import-source on -k regex:k_tile_count -s 1 -c 1 -o gpurun_out/prof_tile_count_r1 -f $CMD > gpurun_out/prof_tile_count_r1.log 2>&1
# the fused GEMM (decoder LSTM step: skip the 33 encoder steps of the first chunk)
ncu --set full --clock-control none --import-source on -k regex:k_gemm_fused -s 40 -c 1 -o gpurun_out/prof_gemm_r1 -f $CMD > gpurun_out/prof_gemm_r1.log 2>&1
ls -la gpurun_out
